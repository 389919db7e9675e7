"""Mirror of halo2_proofs::plonk::permutation::prover::Argument::commit (reference
halo2_proofs/src/plonk/permutation/prover.rs:46-200): the permutation grand-product polynomials z of every column set,
computed on the device from device-resident columns (cqb_permutation_product_dev) and committed with the resident
g_lagrange (params.commit_lagrange, :166), followed by lagrange_to_coeff and coeff_to_extended (:168-171) — the vectors
never leave HBM. The blinding rows come from the caller's rng, as in the reference (:152-155)."""
import ctypes

import numpy as np

from . import _lib
from .fields import R_MOD, fr_from_limbs, fr_to_limbs

FR_DELTA = 0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2  # bn256/fr.rs:87-92 (GENERATOR^(2^S))


def _ptr_array(ptrs):
    return (ctypes.c_void_p * max(len(ptrs), 1))(*[ctypes.c_void_p(p) for p in ptrs])


def product_set_dev(column_ptrs, perm_ptrs, k, beta, gamma, omega, deltaomega, last_z, d_z):
    """one column set (:82-163): z into d_z (2^k Fr on the device); returns the deltaomega of the next set (:144).
    beta/gamma/omega/deltaomega/last_z are canonical Python ints."""
    assert len(column_ptrs) == len(perm_ptrs) and len(column_ptrs) >= 1
    dw = fr_to_limbs(deltaomega).copy()
    _lib.check(_lib.lib().cqb_permutation_product_dev(
        _ptr_array(column_ptrs), _ptr_array(perm_ptrs), len(column_ptrs), k, _lib.p64(fr_to_limbs(beta)), _lib.p64(fr_to_limbs(gamma)),
        _lib.p64(fr_to_limbs(omega)), _lib.p64(fr_to_limbs(FR_DELTA)), _lib.p64(dw), _lib.p64(fr_to_limbs(last_z)), ctypes.c_void_p(d_z)))
    return fr_from_limbs(dw)


def commit_dev(column_ptrs, perm_ptrs, k, cs_degree, blinding_factors, beta, gamma, omega, blind_rows, z_ptrs):
    """Argument::commit over all sets. column_ptrs / perm_ptrs: device pointers of the permutation's columns and of
    pkey.permutations (Lagrange values, 2^k Fr each), in self.columns order; z_ptrs: one 2^k-Fr device buffer per set;
    blind_rows[s]: (blinding_factors, 4) uint64 random Fr for the set's last rows (the reference draws them from its rng).
    Returns last_z after the final set. z buffers hold the Lagrange values the reference commits (:166)."""
    assert cs_degree >= 3, "assert!(pk.vk.cs_degree >= 3)"  # :78
    chunk_len = cs_degree - 2
    n = 1 << k
    lib = _lib.lib()
    deltaomega, last_z = 1, 1
    nsets = (len(column_ptrs) + chunk_len - 1) // chunk_len
    assert len(z_ptrs) == nsets and len(blind_rows) == nsets
    tmp = np.zeros(4, np.uint64)
    for s in range(nsets):
        cols = column_ptrs[s * chunk_len:(s + 1) * chunk_len]
        perms = perm_ptrs[s * chunk_len:(s + 1) * chunk_len]
        deltaomega = product_set_dev(cols, perms, k, beta, gamma, omega, deltaomega, last_z, z_ptrs[s])
        br = np.ascontiguousarray(blind_rows[s], dtype=np.uint64)
        assert br.shape == (blinding_factors, 4)
        if blinding_factors:  # :152-155
            _lib.check(lib.cqb_memcpy_h2d(ctypes.c_void_p(z_ptrs[s] + (n - blinding_factors) * 32), br.ctypes.data_as(ctypes.c_void_p), br.nbytes))
        _lib.check(lib.cqb_memcpy_d2h(tmp.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(z_ptrs[s] + (n - (blinding_factors + 1)) * 32), 32))  # :157
        _lib.check(lib.cqb_sync())
        last_z = fr_from_limbs(tmp)
    return last_z % R_MOD

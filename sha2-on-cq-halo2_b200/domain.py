"""Mirror of halo2_proofs::poly::EvaluationDomain for the hot path (reference halo2_proofs/src/poly/domain.rs).

Constants are computed on the host exactly as EvaluationDomain::new does (domain.rs:39-142); the transforms run on the
device through the C ABI with the surrounding element-wise scalings fused into the NTT passes.
"""
import numpy as np

from . import _lib
from .fields import FR_ROOT_OF_UNITY, FR_S, FR_ZETA, R_MOD, fr_to_limbs


class ExtendedLagrange:
    """Polynomial<F, ExtendedLagrangeCoeff>: values on the zeta-coset of the extended domain. `divided` records a pending
    divide_by_vanishing_poly so that it is fused into extended_to_coeff's first NTT pass (vanishing/prover.rs:84-87)."""

    def __init__(self, values, divided=False):
        self.values = values
        self.divided = divided


class EvaluationDomain:
    def __init__(self, j, k):
        """reference domain.rs:39 EvaluationDomain::new(j, k)"""
        self.quotient_poly_degree = j - 1
        self.k = k
        self.n = 1 << k
        ek = k
        while (1 << ek) < self.n * self.quotient_poly_degree:
            ek += 1
        assert ek <= FR_S
        self.extended_k = ek
        ew = FR_ROOT_OF_UNITY
        for _ in range(ek, FR_S):
            ew = ew * ew % R_MOD
        w = ew
        for _ in range(k, ek):
            w = w * w % R_MOD
        self._omega, self._extended_omega = w, ew
        self._g_coset = FR_ZETA
        self._g_coset_inv = FR_ZETA * FR_ZETA % R_MOD
        orig = pow(FR_ZETA, self.n, R_MOD)
        step = pow(ew, self.n, R_MOD)
        t, cur = [], orig
        while True:
            t.append(cur)
            cur = cur * step % R_MOD
            if cur == orig:
                break
        assert len(t) == 1 << (ek - k)
        self._t_evaluations = [pow((v - 1) % R_MOD, -1, R_MOD) for v in t]
        # Montgomery-limb forms handed to the device
        self.omega = fr_to_limbs(w)
        self.omega_inv = fr_to_limbs(pow(w, -1, R_MOD))
        self.extended_omega = fr_to_limbs(ew)
        self.extended_omega_inv = fr_to_limbs(pow(ew, -1, R_MOD))
        self.g_coset = fr_to_limbs(self._g_coset)
        self.g_coset_inv = fr_to_limbs(self._g_coset_inv)
        self.ifft_divisor = fr_to_limbs(pow(1 << k, -1, R_MOD))
        self.extended_ifft_divisor = fr_to_limbs(pow(1 << ek, -1, R_MOD))
        self.t_evaluations = np.stack([fr_to_limbs(v) for v in self._t_evaluations])

    def extended_len(self):
        return 1 << self.extended_k

    @staticmethod
    def ifft(a, omega_inv, log_n, divisor):
        """reference domain.rs:366-374 (in place)"""
        assert a.shape[0] == 1 << log_n
        _lib.check(_lib.lib().cqb_intt_bn254_fr(_lib.p64(a), _lib.p64(_lib.fr_limbs(omega_inv)), _lib.p64(_lib.fr_limbs(divisor)), log_n))

    def lagrange_to_coeff(self, a):
        """reference domain.rs:238-248: consumes the Lagrange vector, returns coefficients"""
        a = np.array(a, dtype=np.uint64, copy=True)
        assert a.shape == (1 << self.k, 4), "assert_eq!(a.values.len(), 1 << self.k)"
        self.ifft(a, self.omega_inv, self.k, self.ifft_divisor)
        return a

    def coeff_to_extended(self, a):
        """reference domain.rs:252-266"""
        a = np.ascontiguousarray(a, dtype=np.uint64)
        assert a.shape == (1 << self.k, 4), "assert_eq!(a.values.len(), 1 << self.k)"
        out = np.empty((self.extended_len(), 4), np.uint64)
        _lib.check(_lib.lib().cqb_coset_ntt_bn254_fr(_lib.p64(a), a.shape[0], _lib.p64(out), _lib.p64(self.extended_omega),
                                                     self.extended_k, _lib.p64(self.g_coset), _lib.p64(self.g_coset_inv)))
        return ExtendedLagrange(out)

    def divide_by_vanishing_poly(self, a):
        """reference domain.rs:319-338 — recorded, and executed fused with the following extended_to_coeff"""
        assert isinstance(a, ExtendedLagrange) and a.values.shape[0] == self.extended_len()
        assert not a.divided
        return ExtendedLagrange(a.values, divided=True)

    def extended_to_coeff(self, a):
        """reference domain.rs:293-315: returns n * quotient_poly_degree coefficients (truncated as the reference does)"""
        assert isinstance(a, ExtendedLagrange) and a.values.shape[0] == self.extended_len(), \
            "assert_eq!(a.values.len(), self.extended_len())"
        v = np.array(a.values, dtype=np.uint64, copy=True)
        tev = self.t_evaluations if a.divided else None
        _lib.check(_lib.lib().cqb_coset_intt_bn254_fr(
            _lib.p64(v), self.extended_k, _lib.p64(self.extended_omega_inv), _lib.p64(self.extended_ifft_divisor),
            _lib.p64(self.g_coset), _lib.p64(self.g_coset_inv), _lib.p64(tev) if tev is not None else None,
            0 if tev is None else tev.shape[0]))
        return v[: self.n * self.quotient_poly_degree]

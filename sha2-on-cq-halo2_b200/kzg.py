"""Mirror of halo2_proofs::poly::kzg::commitment::ParamsKZG's commit path (reference poly/kzg/commitment.rs:31-39,
496-504, 539-543) and of TableSRS (:42-47) as MSM operands. The SRS vectors are uploaded once and stay resident in HBM
(one `cqb_bases_t` handle each); commits only move the n x 32 B scalars."""
import ctypes

import numpy as np

from . import _lib
from .arithmetic import G1, _as_fr, _as_g1


class DeviceBases:
    """A device-resident Vec<G1Affine>"""

    def __init__(self, affine, precompute=None, window_bits=0):
        """precompute: build the resident 2^(c w) P_i table (cqb_bases_precompute) so that MSMs over this set use one
        bucket set with wider windows. None = automatically for sets of >= 2^16 points."""
        affine = _as_g1(affine)
        self.n = affine.shape[0]
        h = ctypes.c_uint64(0)
        _lib.check(_lib.lib().cqb_bases_register(_lib.p64(affine), self.n, ctypes.byref(h)))
        self.handle = h.value
        if precompute is None:
            precompute = self.n >= (1 << 16)
        if precompute and self.n:
            self.precompute(window_bits)

    def precompute(self, window_bits=0):
        _lib.check(_lib.lib().cqb_bases_precompute(self.handle, window_bits))

    def free(self):
        if self.handle:
            _lib.check(_lib.lib().cqb_bases_free(self.handle))
            self.handle = 0

    def msm(self, scalars, offset=0):
        scalars = _as_fr(scalars)
        out = np.zeros(8, np.uint64)
        inf = ctypes.c_int(0)
        _lib.check(_lib.lib().cqb_msm_bn254_g1(self.handle, offset, _lib.p64(scalars), scalars.shape[0], _lib.p64(out), ctypes.byref(inf)))
        return G1(out, inf.value)

    def msm_sparse(self, idx, scalars):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        scalars = _as_fr(scalars)
        assert idx.shape[0] == scalars.shape[0]
        out = np.zeros(8, np.uint64)
        inf = ctypes.c_int(0)
        _lib.check(_lib.lib().cqb_msm_bn254_g1_sparse(self.handle, idx.ctypes.data_as(_lib.u32p), _lib.p64(scalars), idx.shape[0],
                                                      _lib.p64(out), ctypes.byref(inf)))
        return G1(out, inf.value)


class ParamsKZG:
    """reference poly/kzg/commitment.rs:31-39 { k, n, g, g_lagrange, .. } (g2 / s_g2 are verifier-side, out of scope)"""

    def __init__(self, k, g, g_lagrange, precompute=None):
        self.k = k
        self.n = 1 << k
        assert g.shape == (self.n, 8) and g_lagrange.shape == (self.n, 8)
        self.g = DeviceBases(g, precompute)
        self.g_lagrange = DeviceBases(g_lagrange, precompute)

    def commit_lagrange(self, poly, _blind=None):
        """reference commitment.rs:496-504: assert!(self.n() >= size); best_multiexp(poly, &self.g_lagrange[0..size])"""
        poly = _as_fr(poly)
        assert self.n >= poly.shape[0], "assert!(self.n() >= size as u64)"
        return self.g_lagrange.msm(poly)

    def commit(self, poly, _blind=None):
        """reference commitment.rs:539-543: best_multiexp(poly, &self.g[0..size])"""
        poly = _as_fr(poly)
        assert self.n >= poly.shape[0], "assert!(self.n() >= size as u64)"
        return self.g.msm(poly)

    def free(self):
        self.g.free()
        self.g_lagrange.free()


class TableSRS:
    """reference commitment.rs:42-47 (G1 parts): g1, g1_lagrange, g_lagrange_opening_at_0, device resident"""

    def __init__(self, g1, g1_lagrange, g_lagrange_opening_at_0):
        self.size = g1.shape[0]
        self.g1 = DeviceBases(g1)
        self.g1_lagrange = DeviceBases(g1_lagrange)
        self.g_lagrange_opening_at_0 = DeviceBases(g_lagrange_opening_at_0)

    def free(self):
        self.g1.free()
        self.g1_lagrange.free()
        self.g_lagrange_opening_at_0.free()

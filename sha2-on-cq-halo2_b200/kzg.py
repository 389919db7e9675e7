"""Mirror of halo2_proofs::poly::kzg::commitment::ParamsKZG's commit path (reference poly/kzg/commitment.rs:31-39,
496-504, 539-543) and of TableSRS (:42-47) as MSM operands. The SRS vectors are uploaded once and stay resident in HBM
(one `cqb_bases_t` handle each); commits only move the n x 32 B scalars."""
import ctypes

import numpy as np

from . import _lib
from .arithmetic import G1, _as_fr, _as_g1


# Sets from this size on get their per-SRS table by default: without one an MSM ends in the ~254 dependent doublings of the window
# combination (1.8 ms on one thread), which is most of the time of a commitment at circuit sizes k <= 16; the table of a small set is
# built in microseconds and costs (254/c + 1) x n x 64 bytes
PRECOMPUTE_MIN_POINTS = 1 << 10


class DeviceBases:
    """A device-resident Vec<G1Affine>"""

    @classmethod
    def adopt(cls, device_ptr, n, precompute=None, window_bits=0):
        """wrap points that already live in device memory (no host copy); the caller keeps the allocation alive"""
        self = cls.__new__(cls)
        self.n = n
        h = ctypes.c_uint64(0)
        _lib.check(_lib.lib().cqb_bases_register_device(ctypes.c_void_p(device_ptr), n, ctypes.byref(h)))
        self.handle = h.value
        self._device_ptr = device_ptr
        if precompute is None:
            precompute = n >= PRECOMPUTE_MIN_POINTS
        if precompute and n:
            self.precompute(window_bits)
        return self

    def to_host(self):
        """download the points (ParamsKZG::write_custom needs them back: read -> write and downsize round trips)"""
        out = np.zeros((self.n, 8), np.uint64)
        _lib.check(_lib.lib().cqb_bases_download(self.handle, 0, self.n, _lib.p64(out)))
        return out

    def __init__(self, affine, precompute=None, window_bits=0):
        """precompute: build the resident 2^(c w) P_i table (cqb_bases_precompute) so that MSMs over this set use one
        bucket set with wider windows. None = automatically for sets of >= PRECOMPUTE_MIN_POINTS points."""
        affine = _as_g1(affine)
        self.n = affine.shape[0]
        h = ctypes.c_uint64(0)
        _lib.check(_lib.lib().cqb_bases_register(_lib.p64(affine), self.n, ctypes.byref(h)))
        self.handle = h.value
        if precompute is None:
            precompute = self.n >= PRECOMPUTE_MIN_POINTS
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

    def msm_batch(self, polys, offset=0):
        """commit to several polynomials of equal length at once (plonk/prover.rs:356-360 commits every advice column in a
        loop); polys: (B, n, 4) uint64 -> list of G1"""
        polys = np.ascontiguousarray(polys, dtype=np.uint64)
        assert polys.ndim == 3 and polys.shape[2] == 4
        B, n = polys.shape[0], polys.shape[1]
        out = np.zeros((B, 8), np.uint64)
        inf = (ctypes.c_int * B)()
        _lib.check(_lib.lib().cqb_msm_bn254_g1_batch(self.handle, offset, _lib.p64(polys), n, B, _lib.p64(out), inf))
        return [G1(out[b].copy(), inf[b]) for b in range(B)]

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

    def __init__(self, k, g, g_lagrange, precompute=None, g2=None, s_g2=None):
        self.k = k
        self.n = 1 << k
        assert g.shape == (self.n, 8) and g_lagrange.shape == (self.n, 8)
        self.g = DeviceBases(g, precompute)
        self.g_lagrange = DeviceBases(g_lagrange, precompute)
        self.g2, self.s_g2 = g2, s_g2  # 128-byte raw G2Affine blobs: verifier-side data, carried opaquely
        self._dev_alloc = None

    @classmethod
    def setup_from_toxic_waste(cls, k, s, precompute=None):
        """reference poly/kzg/commitment.rs:209-276, G1 part: g[i] = [s^i]G and g_lagrange[i] = [L_i(s)]G are generated
        ON THE DEVICE (cqb_srs_setup_dev) and registered in place; the SRS never exists in host memory. `s`: (4,) uint64
        Montgomery limbs. (g2 / s_g2 = [s]G2 are verifier-side; supply them separately if the params are to be written.)"""
        assert k <= 28, "assert!(k <= E::Scalar::S)"  # commitment.rs:212
        lib = _lib.lib()
        n = 1 << k
        d = ctypes.c_void_p()
        _lib.check(lib.cqb_dev_alloc(2 * n * 64, ctypes.byref(d)))
        _lib.check(lib.cqb_srs_setup_dev(k, _lib.p64(_lib.fr_limbs(s)), d, ctypes.c_void_p(d.value + n * 64)))
        _lib.check(lib.cqb_sync())
        self = cls.__new__(cls)
        self.k, self.n = k, n
        self.g = DeviceBases.adopt(d.value, n, precompute)
        self.g_lagrange = DeviceBases.adopt(d.value + n * 64, n, precompute)
        # g2 = G2 generator, s_g2 = [s]G2 (:265-266): the first two G2 powers, as the 128-byte raw blobs write() emits
        p2 = g2_powers(s, 2)
        self.g2, self.s_g2 = p2[0].tobytes(), p2[1].tobytes()
        self._dev_alloc = d
        return self

    def downsize(self, k):
        """reference commitment.rs:482-490: truncate g to 2^k points and rebuild g_lagrange with g_to_lagrange (the G1
        EC-FFT of arithmetic.rs:277-301) — on the device, from the resident monomial SRS"""
        assert k <= self.k, "assert!(k <= self.k)"
        lib = _lib.lib()
        n = 1 << k
        d = ctypes.c_void_p()
        _lib.check(lib.cqb_dev_alloc(2 * n * 64, ctypes.byref(d)))
        _lib.check(lib.cqb_bases_copy_dev(self.g.handle, 0, n, d))
        _lib.check(lib.cqb_g_to_lagrange_dev(d, k, ctypes.c_void_p(d.value + n * 64)))
        _lib.check(lib.cqb_sync())
        old = self._dev_alloc
        self.g.free()
        self.g_lagrange.free()
        if old is not None:  # params read from a file / built from host arrays own their points through the handles
            _lib.check(lib.cqb_dev_free(old))
        self.k, self.n, self._dev_alloc = k, n, d
        self.g = DeviceBases.adopt(d.value, n, precompute=False)
        self.g_lagrange = DeviceBases.adopt(d.value + n * 64, n, precompute=False)

    def write(self, writer, g2=None, s_g2=None):
        """reference commitment.rs:366-380 write_custom(RawBytesUnchecked): k as u32 LE, g, g_lagrange (x||y Montgomery
        limbs, 64 B each), then g2 and s_g2 (128 B raw G2Affine each)"""
        g2 = g2 if g2 is not None else self.g2
        s_g2 = s_g2 if s_g2 is not None else self.s_g2
        assert g2 is not None and s_g2 is not None and len(g2) == 128 and len(s_g2) == 128, "g2 / s_g2 raw bytes required"
        writer.write(int(self.k).to_bytes(4, "little"))
        for b in (self.g, self.g_lagrange):
            writer.write(b.to_host().tobytes())
        writer.write(bytes(g2))
        writer.write(bytes(s_g2))

    @classmethod
    def read(cls, reader, precompute=None):
        """reference commitment.rs:383-459 read_custom(RawBytesUnchecked): the point arrays go straight to the device"""
        k = int.from_bytes(reader.read(4), "little")
        n = 1 << k
        g = np.frombuffer(reader.read(n * 64), dtype=np.uint64).reshape(n, 8)
        gl = np.frombuffer(reader.read(n * 64), dtype=np.uint64).reshape(n, 8)
        g2, s_g2 = reader.read(128), reader.read(128)
        assert len(s_g2) == 128, "truncated params file"
        return cls(k, g, gl, precompute, g2=g2, s_g2=s_g2)

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
        if self._dev_alloc is not None:
            _lib.check(_lib.lib().cqb_dev_free(self._dev_alloc))
            self._dev_alloc = None


def g2_powers(s, count):
    """[s^i] G2 for i < count as a (count, 16) uint64 G2Affine array (poly/kzg/commitment.rs:94-104, 114-141; 265-266 for i = 1)"""
    out = np.zeros((count, 16), np.uint64)
    _lib.check(_lib.lib().cqb_g2_powers(_lib.p64(_lib.fr_limbs(s)), count, _lib.p64(out)))
    return out


class MSMKZG:
    """reference poly/kzg/msm.rs:12-80: a multiscalar multiplication collected term by term with PROJECTIVE bases (E::G1); eval()
    normalises them (Curve::batch_normalize) and calls best_multiexp — here one device call does both."""

    def __init__(self):
        self.scalars, self.bases = [], []  # (4,) uint64 Fr ; (12,) uint64 Jacobian x, y, z

    def append_term(self, scalar, point_jacobian):  # :41-44
        self.scalars.append(np.asarray(scalar, dtype=np.uint64).reshape(4))
        self.bases.append(np.asarray(point_jacobian, dtype=np.uint64).reshape(12))

    def add_msm(self, other):  # :46-49
        self.scalars.extend(other.scalars)
        self.bases.extend(other.bases)

    def eval(self):  # :65-70
        n = len(self.scalars)
        sc = np.ascontiguousarray(np.stack(self.scalars)) if n else np.zeros((0, 4), np.uint64)
        bs = np.ascontiguousarray(np.stack(self.bases)) if n else np.zeros((0, 12), np.uint64)
        out = np.zeros(8, np.uint64)
        inf = ctypes.c_int(0)
        _lib.check(_lib.lib().cqb_msm_bn254_g1_jacobian(_lib.p64(bs), _lib.p64(sc), n, _lib.p64(out), ctypes.byref(inf)))
        return G1(out, inf.value)

    def check(self):  # :60-62
        return self.eval().is_identity


def batch_normalize(points_jacobian):
    """Curve::batch_normalize (arithmetic/curves/src/derive/curve.rs:362-397): (n, 12) Jacobian -> (n, 8) affine, identities -> zeros"""
    p = np.ascontiguousarray(points_jacobian, dtype=np.uint64)
    assert p.ndim == 2 and p.shape[1] == 12
    out = np.zeros((p.shape[0], 8), np.uint64)
    _lib.check(_lib.lib().cqb_g1_batch_normalize(_lib.p64(p), p.shape[0], _lib.p64(out)))
    return out


class TableSRS:
    """reference commitment.rs:42-47 (G1 parts): g1, g1_lagrange, g_lagrange_opening_at_0, device resident"""

    def __init__(self, g1, g1_lagrange, g_lagrange_opening_at_0):
        self.size = g1.shape[0]
        self.g1 = DeviceBases(g1)
        self.g1_lagrange = DeviceBases(g1_lagrange)
        self.g_lagrange_opening_at_0 = DeviceBases(g_lagrange_opening_at_0)
        self._dev_alloc = None

    @classmethod
    def setup_from_toxic_waste(cls, max_g1_power, s, precompute=None, max_g2_power=None):
        """reference commitment.rs:73-178: generated on the device. max_g2_power: also build g2 = [s^i]G2, i <= max_g2_power (:94-104),
        as a host (count, 16) uint64 array — keygen / verifier-side data (StaticTableValues.commit reads it)"""
        g1_len = max_g1_power + 1
        assert g1_len & (g1_len - 1) == 0, "assert!(is_pow_2(g1_len))"  # commitment.rs:77
        log_len = g1_len.bit_length() - 1
        lib = _lib.lib()
        d = ctypes.c_void_p()
        _lib.check(lib.cqb_dev_alloc(3 * g1_len * 64, ctypes.byref(d)))
        _lib.check(lib.cqb_table_srs_setup_dev(log_len, _lib.p64(_lib.fr_limbs(s)), d, ctypes.c_void_p(d.value + g1_len * 64),
                                               ctypes.c_void_p(d.value + 2 * g1_len * 64)))
        _lib.check(lib.cqb_sync())
        self = cls.__new__(cls)
        self.size = g1_len
        self.g1 = DeviceBases.adopt(d.value, g1_len, precompute)
        self.g1_lagrange = DeviceBases.adopt(d.value + g1_len * 64, g1_len, precompute)
        self.g_lagrange_opening_at_0 = DeviceBases.adopt(d.value + 2 * g1_len * 64, g1_len, precompute)
        self.g2 = g2_powers(s, max_g2_power + 1) if max_g2_power is not None else None
        self._dev_alloc = d
        return self

    def free(self):
        self.g1.free()
        self.g1_lagrange.free()
        self.g_lagrange_opening_at_0.free()
        if self._dev_alloc is not None:
            _lib.check(_lib.lib().cqb_dev_free(self._dev_alloc))
            self._dev_alloc = None

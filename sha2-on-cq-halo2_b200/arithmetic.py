"""Mirror of halo2_proofs::arithmetic for the hot path (reference halo2_proofs/src/arithmetic.rs).

    best_multiexp(coeffs, bases) -> G1      arithmetic.rs:132-159
    best_fft(a, omega, log_n)               arithmetic.rs:171-234   (in place)

Same names, argument meaning and error behaviour (length mismatch / wrong size raise AssertionError, the Python
counterpart of the reference's assert_eq! panics). Arrays are numpy uint64 in the reference's in-memory layout:
Fr = (...,4) Montgomery limbs, G1Affine = (...,8) x||y with identity = zeros.
"""
import ctypes

import numpy as np

from . import _lib
from .fields import FQ_ONE_MONT


class G1:
    """C::Curve as the reference returns it from best_multiexp: projective (x, y, z). Only the affine normal form is
    canonical, so the device returns that and z is one (or zero for the identity)."""

    __slots__ = ("affine", "is_identity")

    def __init__(self, affine, is_identity):
        self.affine = affine
        self.is_identity = bool(is_identity)

    def to_affine(self):
        """group::Curve::to_affine (derive/curve.rs:399-412): (8,) uint64 x||y, zeros for the identity"""
        return self.affine

    def jacobian_coordinates(self):
        z = np.zeros(4, np.uint64) if self.is_identity else np.array(
            [(FQ_ONE_MONT >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
        return self.affine[:4].copy(), self.affine[4:].copy(), z

    def __eq__(self, other):
        return isinstance(other, G1) and np.array_equal(self.affine, other.affine)

    def __repr__(self):
        return "G1(identity)" if self.is_identity else f"G1(x={[hex(int(v)) for v in self.affine[:4]]}, ...)"


def _as_fr(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.ndim == 2 and a.shape[1] == 4, f"expected (n,4) uint64 Fr array, got {a.shape}"
    return a


def _as_g1(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.ndim == 2 and a.shape[1] == 8, f"expected (n,8) uint64 G1Affine array, got {a.shape}"
    return a


def best_multiexp(coeffs, bases):
    """reference arithmetic.rs:132: "This function will panic if coeffs and bases have a different length." """
    coeffs = _as_fr(coeffs)
    bases = _as_g1(bases)
    assert coeffs.shape[0] == bases.shape[0], "assert_eq!(coeffs.len(), bases.len())"  # arithmetic.rs:133
    lib = _lib.lib()
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    _lib.check(lib.cqb_msm_bn254_g1_host(_lib.p64(bases), _lib.p64(coeffs), coeffs.shape[0], _lib.p64(out), ctypes.byref(inf)))
    return G1(out, inf.value)


def best_fft(a, omega, log_n):
    """reference arithmetic.rs:171: in place; a must be a writable contiguous (2^log_n, 4) uint64 array"""
    assert isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"]
    assert a.ndim == 2 and a.shape[1] == 4
    assert a.shape[0] == 1 << log_n, "assert_eq!(n, 1 << log_n)"  # arithmetic.rs:184
    lib = _lib.lib()
    _lib.check(lib.cqb_ntt_bn254_fr(_lib.p64(a), _lib.p64(_lib.fr_limbs(omega)), log_n))


def _with_device_copy(a):
    lib = _lib.lib()
    d = ctypes.c_void_p()
    _lib.check(lib.cqb_dev_alloc(max(a.nbytes, 64), ctypes.byref(d)))
    if a.nbytes:
        _lib.check(lib.cqb_memcpy_h2d(d, a.ctypes.data_as(ctypes.c_void_p), a.nbytes))
    return d


def eval_polynomial(poly, point):
    """reference arithmetic.rs:304-329: evaluates a polynomial in coefficient form at `point` -> (4,) uint64"""
    poly = _as_fr(poly)
    lib = _lib.lib()
    d = _with_device_copy(poly)
    out = np.zeros(4, np.uint64)
    try:
        _lib.check(lib.cqb_eval_polynomial_dev(d, poly.shape[0], _lib.p64(_lib.fr_limbs(point)), _lib.p64(out)))
    finally:
        _lib.check(lib.cqb_dev_free(d))
    return out


def kate_division(a, b):
    """reference arithmetic.rs:351-387: divides a(X) by (X - b) with no remainder -> (n-1, 4) uint64"""
    a = _as_fr(a)
    n = a.shape[0]
    assert n >= 1
    lib = _lib.lib()
    q = np.zeros((n - 1, 4), np.uint64)
    if n == 1:
        return q
    d = _with_device_copy(a)
    dq = ctypes.c_void_p()
    _lib.check(lib.cqb_dev_alloc(q.nbytes, ctypes.byref(dq)))
    try:
        _lib.check(lib.cqb_kate_division_dev(d, n, _lib.p64(_lib.fr_limbs(b)), dq))
        _lib.check(lib.cqb_memcpy_d2h(q.ctypes.data_as(ctypes.c_void_p), dq, q.nbytes))
    finally:
        _lib.check(lib.cqb_dev_free(d))
        _lib.check(lib.cqb_dev_free(dq))
    return q

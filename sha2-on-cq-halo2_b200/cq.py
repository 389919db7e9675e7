"""The CQ (cached quotients) static-lookup prover's commitment calls (reference plonk/static_lookup/prover.rs).

The reference commits to the sparse polynomials m(X), A(X), Q_A(X), A_0(X) with a serial loop of 256-bit scalar
multiplications over the support of m (prover.rs:167-170, 245-257). Here each is ONE sparse MSM on the device that
returns the same group element. Dense commitments (f, B_0, P) go through ParamsKZG / best_multiexp
exactly as in the reference (prover.rs:165, 299, 310).
"""
import numpy as np

from .kzg import DeviceBases


def commit_m(table_srs, m_sparse):
    """prover.rs:167-170: m_cm = sum_{(index, multiplicity)} g1_lagrange[index] * multiplicity.
    m_sparse: dict {index: (4,) uint64 Fr} (the reference's BTreeMap<usize, Scalar>)"""
    idx = np.array(sorted(m_sparse.keys()), dtype=np.uint32)
    sc = np.stack([np.asarray(m_sparse[int(i)], dtype=np.uint64) for i in idx]) if len(idx) else np.zeros((0, 4), np.uint64)
    return table_srs.g1_lagrange.msm_sparse(idx, sc)


def commit_log_derivative_sparse(table_srs, qs_bases, idx, a_values):
    """prover.rs:245-257: (a_cm, qa_cm, a0_cm) for A's sparse values a_i over the support `idx`.
    qs_bases: DeviceBases of the theta-compressed cached quotient commitments (affine), indexed like the table."""
    assert isinstance(qs_bases, DeviceBases)
    a_cm = table_srs.g1_lagrange.msm_sparse(idx, a_values)
    qa_cm = qs_bases.msm_sparse(idx, a_values)
    a0_cm = table_srs.g_lagrange_opening_at_0.msm_sparse(idx, a_values)
    return a_cm, qa_cm, a0_cm


def commit_b0_and_p(params, b0_bound_bases, b_coeffs):
    """prover.rs:279-311: B_0 = (B - B(0))/X; p_cm = best_multiexp(b0[..n-1], pk.b0_g1_bound); b0_cm = params.commit(b0)"""
    b0 = np.ascontiguousarray(b_coeffs[1:], dtype=np.uint64)
    assert b0.shape[0] == b0_bound_bases.n, "assert_eq!(coeffs.len(), bases.len())"  # arithmetic.rs:133 via prover.rs:299
    p_cm = b0_bound_bases.msm(b0)
    b0_full = np.concatenate([b0, np.zeros((1, 4), np.uint64)])
    b0_cm = params.commit(b0_full)
    return b0_cm, p_cm


class StaticTableValues:
    """reference plonk/static_lookup.rs:69-126: the per-table cached quotient commitments `qs` of the CQ argument.

    The reference computes them with N kate_divisions and N MSMs of N-1 points (O(N^2); ":107 TODO: THIS SHOULD BE DONE
    WITH FK METHOD"). Here: ifft of the values on the device, then the FK algorithm (cqb_cq_table_qs_dev: three G1
    EC-NTTs), O(N log N). `qs` stays resident as a DeviceBases so that Q_A commitments (prover.rs:245-257) index it."""

    def __init__(self, values, srs_g1):
        import ctypes

        from . import _lib
        from .domain import EvaluationDomain

        values = np.ascontiguousarray(values, dtype=np.uint64)
        size = values.shape[0]
        assert size & (size - 1) == 0, "assert!(is_pow_2(size))"          # static_lookup.rs:80
        assert len({bytes(v) for v in values}) == size, "table is all unique values"  # :82-85
        assert isinstance(srs_g1, DeviceBases) and srs_g1.n >= size and hasattr(srs_g1, "_device_ptr"), \
            "srs_g1 must be a device-resident SRS with at least `size` powers"
        self.size = size
        self.value_index_mapping = {bytes(v): i for i, v in enumerate(values)}
        k = size.bit_length() - 1
        lib = _lib.lib()
        dom = EvaluationDomain(2, k)
        d_vals = ctypes.c_void_p()
        _lib.check(lib.cqb_dev_alloc(size * 32 + size * 64, ctypes.byref(d_vals)))
        d_qs = ctypes.c_void_p(d_vals.value + size * 32)
        _lib.check(lib.cqb_memcpy_h2d(d_vals, values.ctypes.data_as(ctypes.c_void_p), size * 32))
        _lib.check(lib.cqb_intt_bn254_fr_dev(d_vals, _lib.p64(dom.omega_inv), _lib.p64(dom.ifft_divisor), k))  # :99-105
        _lib.check(lib.cqb_cq_table_qs_dev(d_vals, k, ctypes.c_void_p(srs_g1._device_ptr), d_qs))
        _lib.check(lib.cqb_sync())
        self._dev_alloc = d_vals
        self.qs = DeviceBases.adopt(d_qs.value, size, precompute=False)

    def free(self):
        from . import _lib

        self.qs.free()
        _lib.check(_lib.lib().cqb_dev_free(self._dev_alloc))

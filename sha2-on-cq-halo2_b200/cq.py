"""The CQ (cached quotients) static-lookup prover's commitment calls (reference plonk/static_lookup/prover.rs).

The reference commits to the sparse polynomials m(X), A(X), Q_A(X), A_0(X) with a serial loop of 256-bit scalar
multiplications over the support of m (prover.rs:167-170, 245-257). Here each is ONE sparse MSM on the device that
returns the same group element. Dense commitments (f, B_0, P) go through ParamsKZG / best_multiexp
exactly as in the reference (prover.rs:165, 299, 310).
"""
import ctypes

import numpy as np

from . import _lib
from .arithmetic import G1
from .fields import R_MOD, fr_from_limbs, fr_to_limbs
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
        self._values = values.copy()
        k = size.bit_length() - 1
        lib = _lib.lib()
        dom = EvaluationDomain(2, k)
        d_vals = ctypes.c_void_p()
        _lib.check(lib.cqb_dev_alloc(size * 32 + size * 64 + size * 32, ctypes.byref(d_vals)))
        d_qs = ctypes.c_void_p(d_vals.value + size * 32)
        self.d_values = d_vals.value + size * 32 + size * 64  # the table's values (Lagrange form) stay resident for the prover
        _lib.check(lib.cqb_memcpy_h2d(d_vals, values.ctypes.data_as(ctypes.c_void_p), size * 32))
        _lib.check(lib.cqb_memcpy_d2d(ctypes.c_void_p(self.d_values), d_vals, size * 32))
        _lib.check(lib.cqb_intt_bn254_fr_dev(d_vals, _lib.p64(dom.omega_inv), _lib.p64(dom.ifft_divisor), k))  # :99-105
        _lib.check(lib.cqb_cq_table_qs_dev(d_vals, k, ctypes.c_void_p(srs_g1._device_ptr), d_qs))
        _lib.check(lib.cqb_sync())
        self._dev_alloc = d_vals
        self.qs = DeviceBases.adopt(d_qs.value, size)  # table by default from 2^10 rows on: Q_A is a short sparse MSM over it every proof

    def commit(self, srs_g1_len, srs_g2, circuit_domain):
        """reference plonk/static_lookup.rs:127-160 StaticTableValues::commit -> StaticCommittedTable { zv, t, x_b0_bound, size }:
        zv = [s^N]G2 - G2 (:137); t = best_multiexp::<G2Affine>(ifft(table values), srs_g2) (:139-146) — the values are taken in the
        order of value_index_mapping.keys(), a BTreeMap, i.e. SORTED by canonical value (derive/field.rs:128-141), as the reference
        does; x_b0_bound = srs_g2[srs_g1_len - 1 - (circuit_domain - 2)] (:149). srs_g2: (>= N + 1, 16) uint64 G2Affine array."""
        from . import _lib
        from .domain import EvaluationDomain
        from .fields import fr_from_limbs, fr_to_limbs

        srs_g2 = np.ascontiguousarray(srs_g2, dtype=np.uint64)
        N = self.size
        assert srs_g2.ndim == 2 and srs_g2.shape[1] == 16 and srs_g2.shape[0] > N, "srs_g2 must hold the powers 0..N"
        lib = _lib.lib()
        out = np.zeros(16, np.uint64)
        inf = ctypes.c_int(0)
        pm = np.stack([fr_to_limbs(1), fr_to_limbs(R_MOD - 1)])
        _lib.check(lib.cqb_msm_bn254_g2(_lib.p64(np.ascontiguousarray(srs_g2[[N, 0]])), _lib.p64(pm), 2, _lib.p64(out), ctypes.byref(inf)))
        zv = out.copy()
        order = sorted(range(N), key=lambda i: fr_from_limbs(self._values[i]))
        coeffs = np.ascontiguousarray(self._values[order])
        dom = EvaluationDomain(2, N.bit_length() - 1)
        EvaluationDomain.ifft(coeffs, dom.omega_inv, dom.k, dom.ifft_divisor)
        _lib.check(lib.cqb_msm_bn254_g2(_lib.p64(np.ascontiguousarray(srs_g2[:N])), _lib.p64(coeffs), N, _lib.p64(out), ctypes.byref(inf)))
        return {"zv": zv, "t": out.copy(), "x_b0_bound": srs_g2[srs_g1_len - 1 - (circuit_domain - 2)].copy(), "size": srs_g1_len}

    def free(self):
        from . import _lib

        self.qs.free()
        _lib.check(_lib.lib().cqb_dev_free(self._dev_alloc))


class CommittedLogDerivative:
    """reference static_lookup/prover.rs:36-41 { b, b0, f, a_at_zero } (coefficient form, here device-resident: pointers
    into one allocation), plus the five commitments in the order they are written to the transcript (:301-313)"""

    def __init__(self, alloc, d_b, d_b0, d_f, a_at_zero, a_cm, qa_cm, a0_cm, b0_cm, p_cm):
        self._alloc, self.d_b, self.d_b0, self.d_f, self.a_at_zero = alloc, d_b, d_b0, d_f, a_at_zero
        self.a_cm, self.qa_cm, self.a0_cm, self.b0_cm, self.p_cm = a_cm, qa_cm, a0_cm, b0_cm, p_cm

    def free(self):
        if self._alloc is not None:
            _lib.check(_lib.lib().cqb_dev_free(self._alloc))
            self._alloc = None


def commit_log_derivatives_dev(params, table_srs, tables, b0_bound_bases, k, blinding_factors, d_f, idx, multiplicities, beta, theta, alloc=None):
    """Committed::commit_log_derivatives (static_lookup/prover.rs:187-342) with every vector resident in HBM.

    params: ParamsKZG; table_srs: TableSRS (the table_config of :213-216); tables: the lookup's StaticTableValues (same size,
    :82-84), in table_ids order; b0_bound_bases: pk.b0_g1_bound as DeviceBases (n - 1 points, :299); d_f: device pointer of
    the compressed input expression f in Lagrange form (2^k Fr); idx / multiplicities: m_sparse in key order (uint32 / (m,4)
    Fr); beta, theta: canonical ints. Only the 64-byte commitments and B(0) come back to the host."""
    lib = _lib.lib()
    n = 1 << k
    m = int(len(idx))
    K = len(tables)
    N = tables[0].size
    assert all(t.size == N for t in tables), "Tables should all be of the same size"  # :82-84
    assert b0_bound_bases.n == n - 1, "assert_eq!(coeffs.len(), bases.len())"          # arithmetic.rs:133 via :299
    usable = n - (blinding_factors + 1)                                                  # :259-260
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    mult = np.ascontiguousarray(multiplicities, dtype=np.uint64).reshape(m, 4)
    # one allocation: b (n) | b0 (n) | f coeff (n) | a (m) | tv (m) | mult (m) | idx (m u32)
    nbytes = 3 * n * 32 + 3 * max(m, 1) * 32 + max(m, 1) * 4 + 64
    if alloc is not None:  # the caller's arena (a proof's pooled working memory): nothing to free here
        base, owned = alloc(nbytes), None
    else:
        owned = ctypes.c_void_p()
        _lib.check(lib.cqb_dev_alloc(nbytes, ctypes.byref(owned)))
        base = owned.value
    d_b, d_b0, d_fc = base, base + n * 32, base + 2 * n * 32
    d_a = base + 3 * n * 32
    d_tv, d_mult = d_a + max(m, 1) * 32, d_a + 2 * max(m, 1) * 32
    d_idx = d_a + 3 * max(m, 1) * 32
    vp = ctypes.c_void_p
    beta_l, theta_l = fr_to_limbs(beta), fr_to_limbs(theta)
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)

    def sparse(bases):
        _lib.check(lib.cqb_msm_bn254_g1_sparse_dev(bases.handle, vp(d_idx), vp(d_a), m, _lib.p64(out), ctypes.byref(inf)))
        return G1(out.copy(), inf.value)

    if m:
        _lib.check(lib.cqb_memcpy_h2d(vp(d_idx), idx.ctypes.data_as(vp), m * 4))
        _lib.check(lib.cqb_memcpy_h2d(vp(d_mult), mult.ctypes.data_as(vp), m * 32))
        # :224-229 table_values = fold(values * theta + table value at index); :243 a_i = multiplicity / (table_values + beta)
        tv_ptrs = (ctypes.c_void_p * K)(*[vp(t.d_values) for t in tables])
        _lib.check(lib.cqb_fr_compress_dev(tv_ptrs, K, vp(d_idx), m, _lib.p64(theta_l), vp(d_tv)))
        _lib.check(lib.cqb_fr_inv_shifted_dev(vp(d_tv), m, m, _lib.p64(beta_l), vp(d_a)))
        _lib.check(lib.cqb_fr_mul_dev(vp(d_a), vp(d_mult), m, vp(d_a)))
    # :245-257 a_cm, qa_cm, a0_cm. The reference compresses the cached quotients per index, qs = fold(qs * theta + table.qs[i])
    # (:230-233); the sum over the support is linear, so Q_A = sum_k theta^(K-1-k) * MSM(table_k.qs, a) — K sparse MSMs and a
    # K-term combination instead of |supp| x K scalar multiplications and affine conversions.
    a_cm = sparse(table_srs.g1_lagrange)
    if K == 1:
        qa_cm = sparse(tables[0].qs)
    else:
        # the theta powers go into the scalars (a copy of a in the d_tv scratch, which is free once a exists), the K partial points are
        # added on the device: no K-term windowed MSM (whose window combination alone is ~254 dependent doublings)
        parts = []
        for j, t in enumerate(tables):
            pw = pow(theta, K - 1 - j, R_MOD)
            if pw == 1:
                parts.append(sparse(t.qs).to_affine())
                continue
            _lib.check(lib.cqb_memcpy_d2d(vp(d_tv), vp(d_a), m * 32))
            _lib.check(lib.cqb_fr_scale_dev(vp(d_tv), m, _lib.p64(fr_to_limbs(pw))))
            _lib.check(lib.cqb_msm_bn254_g1_sparse_dev(t.qs.handle, vp(d_idx), vp(d_tv), m, _lib.p64(out), ctypes.byref(inf)))
            parts.append(out.copy())
        pts = np.ascontiguousarray(np.stack(parts))
        _lib.check(lib.cqb_g1_sum_affine(_lib.p64(pts), K, _lib.p64(out), ctypes.byref(inf)))
        qa_cm = G1(out.copy(), inf.value)
    a0_cm = sparse(table_srs.g_lagrange_opening_at_0)
    # :261-276 bs = 1/(f_i + beta) on the usable rows, 1/beta on the blinding rows; ifft
    _lib.check(lib.cqb_fr_inv_shifted_dev(vp(d_f), n, usable, _lib.p64(beta_l), vp(d_b)))
    from .domain import EvaluationDomain

    dom = EvaluationDomain(3, k)
    _lib.check(lib.cqb_intt_bn254_fr_dev(vp(d_b), _lib.p64(dom.omega_inv), _lib.p64(dom.ifft_divisor), k))
    # :279 b0 = bs[1..]; :299 p_cm over the degree-bound bases; :303-304 b0 poly = b0 || 0; :310 b0_cm = params.commit(b0)
    _lib.check(lib.cqb_memcpy_d2d(vp(d_b0), vp(d_b + 32), (n - 1) * 32))
    zero = np.zeros(4, np.uint64)
    _lib.check(lib.cqb_memcpy_h2d(vp(d_b0 + (n - 1) * 32), zero.ctypes.data_as(vp), 32))
    _lib.check(lib.cqb_msm_bn254_g1_dev(b0_bound_bases.handle, 0, vp(d_b0), n - 1, _lib.p64(out), ctypes.byref(inf)))
    p_cm = G1(out.copy(), inf.value)
    _lib.check(lib.cqb_msm_bn254_g1_dev(params.g.handle, 0, vp(d_b0), n, _lib.p64(out), ctypes.byref(inf)))
    b0_cm = G1(out.copy(), inf.value)
    # :315-325 sumcheck identity n B(0) = N A(0): b_at_zero = eval_polynomial(b, 0) = b's constant coefficient
    b0_limbs = np.zeros(4, np.uint64)
    _lib.check(lib.cqb_memcpy_d2h(b0_limbs.ctypes.data_as(vp), vp(d_b), 32))
    _lib.check(lib.cqb_sync())
    b_at_zero = fr_from_limbs(b0_limbs)
    beta_inv = pow(beta, -1, R_MOD)
    a_at_zero = (b_at_zero * n - (blinding_factors + 1) * beta_inv) * pow(N, -1, R_MOD) % R_MOD
    # :327-334 f -> coefficient form
    _lib.check(lib.cqb_memcpy_d2d(vp(d_fc), vp(d_f), n * 32))
    _lib.check(lib.cqb_intt_bn254_fr_dev(vp(d_fc), _lib.p64(dom.omega_inv), _lib.p64(dom.ifft_divisor), k))
    _lib.check(lib.cqb_sync())
    return CommittedLogDerivative(owned, d_b, d_b0, d_fc, a_at_zero, a_cm, qa_cm, a0_cm, b0_cm, p_cm)

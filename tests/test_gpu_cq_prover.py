"""Committed::commit_log_derivatives (reference plonk/static_lookup/prover.rs:187-342) with every vector resident in HBM
(cq.commit_log_derivatives_dev) against the reference's own loop restated over the oracle's primitives: per-index theta-
compression of table values AND cached quotient commitments (:220-240), the serial scalar-multiplication loop (:242-257),
bs / ifft / B_0 / P (:259-311), the sumcheck value A(0) (:315-325), f in coefficient form (:327-334). Two tables (vector
lookup), tables built on the device with the FK preprocessing (StaticTableValues)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402


@pytest.fixture(scope="module")
def cq():
    import cqb200

    cqb200._lib.init(0)
    return cqb200


def F(vals):
    return P.fr_array_from_ints(vals)


def I(arr):
    return P.fr_array_to_ints(np.asarray(arr, dtype=np.uint64).reshape(-1, 4))


@pytest.mark.parametrize("k,N,ntables", [(5, 32, 2), (8, 64, 2), (8, 64, 1), (6, 128, 3)])
def test_commit_log_derivatives_device_resident(cq, oracle, k, N, ntables):
    O = oracle
    L = cq._lib
    lib = L.lib()
    n = 1 << k
    bf = 5
    usable = n - (bf + 1)
    rng = np.random.default_rng(1000 + k + N)
    s = O.synth_scalars(0xC9 + k, 1)[0]
    Nt = max(N, n)
    # ---- keygen-side material: params, table SRS (device-generated), tables with their cached quotients (FK on the device)
    g, g_lagrange = O.params_setup(k, s)
    params = cq.ParamsKZG(k, g, g_lagrange)
    table_srs = cq.TableSRS.setup_from_toxic_waste(Nt - 1, s, precompute=False)
    t_g1, t_lag, t_op0 = (b.to_host() for b in (table_srs.g1, table_srs.g1_lagrange, table_srs.g_lagrange_opening_at_0))
    if Nt == N:
        tsrs = table_srs
    else:  # the table's own SRS (size N) for its Lagrange bases; the big one only supplies the degree-bound slice
        tsrs = cq.TableSRS.setup_from_toxic_waste(N - 1, s, precompute=False)
        t_lag, t_op0 = tsrs.g1_lagrange.to_host(), tsrs.g_lagrange_opening_at_0.to_host()
    exp_g1, exp_lag, exp_op0 = O.table_srs_setup(N, s)
    assert np.array_equal(t_lag, exp_lag) and np.array_equal(t_op0, exp_op0)
    tvals = [[int(v) + (j << 40) for v in rng.choice(1 << 30, N, replace=False)] for j in range(ntables)]
    tables = [cq.cq.StaticTableValues(F(v), tsrs.g1) for v in tvals]
    qs_host = [O.cq_table_qs(F(v), exp_g1, 4) for v in tvals]  # reference static_lookup.rs:77-126 (O(N^2) loop)
    for t, q in zip(tables, qs_host):
        assert np.array_equal(t.qs.to_host(), q)
    b0_bound = t_g1[Nt - (n - 1):]          # my_test.rs:205: the last n-1 powers of the table SRS
    b0_bound_dev = cq.DeviceBases(b0_bound)
    # ---- witness: rows look up row r of every table (vector lookup); f = theta-compression of the inputs (:108-117)
    theta, beta = 0x123456789ABCDEF, 0xFEDCBA9876543
    rows = [int(v) for v in rng.integers(0, N, usable)]
    inputs = [[tv[r] for r in rows] + [int(v) for v in rng.integers(0, 1 << 50, n - usable)] for tv in tvals]
    d_inputs = []
    for col in inputs:
        d = ctypes.c_void_p()
        L.check(lib.cqb_dev_alloc(n * 32, ctypes.byref(d)))
        a = F(col)
        L.check(lib.cqb_memcpy_h2d(d, a.ctypes.data_as(ctypes.c_void_p), n * 32))
        d_inputs.append(d)
    d_f = ctypes.c_void_p()
    L.check(lib.cqb_dev_alloc(n * 32, ctypes.byref(d_f)))
    ptrs = (ctypes.c_void_p * ntables)(*d_inputs)
    L.check(lib.cqb_fr_compress_dev(ptrs, ntables, None, n, L.p64(F([theta])[0]), d_f))
    f_int = [0] * n
    for col in inputs:
        f_int = [(a * theta + c) % P.R_MOD for a, c in zip(f_int, col)]
    f_host = np.zeros((n, 4), np.uint64)
    L.check(lib.cqb_memcpy_d2h(f_host.ctypes.data_as(ctypes.c_void_p), d_f, n * 32))
    L.check(lib.cqb_sync())
    assert I(f_host) == f_int
    m = {}
    for r in rows:                          # :132-160 m_sparse (BTreeMap: key order)
        m[r] = m.get(r, 0) + 1
    idx = np.array(sorted(m), dtype=np.uint32)
    mult = F([m[int(i)] for i in idx])

    # ---- the reference's loop over the oracle's primitives ---------------------------------------------------------------
    one = lambda x: F([x])[0]  # noqa: E731
    a_acc = qa_acc = a0_acc = None
    jac_add = lambda acc, p: p if acc is None else O.g1_add_jj(acc, p)  # noqa: E731
    for i in idx:
        i = int(i)
        values, qs = 0, None
        for j in range(ntables):            # compress_tables (:220-240)
            values = (values * theta + tvals[j][i]) % P.R_MOD
            scaled = O.g1_mul_a(qs, one(theta)) if qs is not None else None   # qs * theta (identity * theta = identity)
            nxt = O.g1_to_curve(qs_host[j][i]) if scaled is None else O.g1_add_ja(scaled, qs_host[j][i])
            qs = O.g1_to_affine(nxt)        # `.into()` affine (:236)
        a_i = m[i] * pow((values + beta) % P.R_MOD, -1, P.R_MOD) % P.R_MOD
        a_acc = jac_add(a_acc, O.g1_mul_a(exp_lag[i], one(a_i)))
        qa_acc = jac_add(qa_acc, O.g1_mul_a(qs, one(a_i)))
        a0_acc = jac_add(a0_acc, O.g1_mul_a(exp_op0[i], one(a_i)))
    exp_a, exp_qa, exp_a0 = (O.g1_to_affine(x) for x in (a_acc, qa_acc, a0_acc))
    beta_inv = pow(beta, -1, P.R_MOD)
    bs = F([pow((fv + beta) % P.R_MOD, -1, P.R_MOD) for fv in f_int[:usable]] + [beta_inv] * (bf + 1))
    odom = O.domain_new(3, k)
    b_coeff = O.lagrange_to_coeff(odom, bs)
    b0 = np.ascontiguousarray(b_coeff[1:])
    exp_p = O.best_multiexp(b0, b0_bound, 2)[1]
    b0_full = np.concatenate([b0, np.zeros((1, 4), np.uint64)])
    exp_b0 = O.best_multiexp(b0_full, g, 2)[1]
    b_at_zero = I(b_coeff[:1])[0]
    exp_a_at_zero = (b_at_zero * n - (bf + 1) * beta_inv) * pow(N, -1, P.R_MOD) % P.R_MOD
    exp_f_coeff = O.lagrange_to_coeff(odom, F(f_int))

    # ---- device ------------------------------------------------------------------------------------------------------------
    got = cq.cq.commit_log_derivatives_dev(params, tsrs, tables, b0_bound_dev, k, bf, d_f.value, idx, mult, beta, theta)
    assert np.array_equal(got.a_cm.to_affine(), exp_a)
    assert np.array_equal(got.qa_cm.to_affine(), exp_qa)
    assert np.array_equal(got.a0_cm.to_affine(), exp_a0)
    assert np.array_equal(got.p_cm.to_affine(), exp_p)
    assert np.array_equal(got.b0_cm.to_affine(), exp_b0)
    assert got.a_at_zero == exp_a_at_zero

    def fetch(ptr):
        out = np.zeros((n, 4), np.uint64)
        L.check(lib.cqb_memcpy_d2h(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(ptr), n * 32))
        L.check(lib.cqb_sync())
        return out

    assert np.array_equal(fetch(got.d_b), b_coeff)
    assert np.array_equal(fetch(got.d_b0), b0_full)
    assert np.array_equal(fetch(got.d_f), exp_f_coeff)
    # the sumcheck identity the verifier relies on: N * A(0) = n * B(0) - (blinding rows) / beta, with A(0) = sum a_i / N
    a_sum = sum(m[int(i)] * pow((sum(tvals[j][int(i)] * pow(theta, ntables - 1 - j, P.R_MOD) for j in range(ntables)) + beta) % P.R_MOD, -1, P.R_MOD)
                for i in idx) % P.R_MOD
    assert a_sum * pow(N, -1, P.R_MOD) % P.R_MOD == exp_a_at_zero
    got.free()
    for d in d_inputs + [d_f]:
        L.check(lib.cqb_dev_free(d))
    for t in tables:
        t.free()
    b0_bound_dev.free()
    params.free()
    table_srs.free()
    if tsrs is not table_srs:
        tsrs.free()

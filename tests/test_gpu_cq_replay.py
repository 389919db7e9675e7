"""BASELINE.json configs[0]: the only live CQ prove+verify in the reference is halo2_proofs/tests/my_test.rs::my_test_e2e
(k=3, two advice columns, one lookup_static over two 16-entry tables). The proof bytes are not
reproducible (OsRng blinds, no golden bytes), so this test replays that test's MSM / NTT call sequence through the C ABI —
same sizes, same call order as plonk/prover.rs:51 -> static_lookup/prover.rs:51,187 -> vanishing/prover.rs:69 — and
compares every commitment (affine + compressed bytes) and every polynomial with the CPU oracle. A scaled copy (k=10,
table 256) checks the same shape at a size where the kernels are not degenerate."""
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


@pytest.mark.parametrize("k,N", [(3, 16), (10, 256)])
def test_cq_commit_sequence_replay(cq, oracle, k, N):
    O = oracle
    n = 1 << k
    rng = np.random.default_rng(k)
    blinding_factors = 5
    usable = n - (blinding_factors + 1)  # static_lookup/prover.rs:129-130
    s_int = P.fr_array_to_ints(O.synth_scalars(0xC9, 1))[0]
    s = P.fr_array_from_ints([s_int])[0]
    # SRS: circuit params (k) and the table SRS (N), my_test.rs:182-210
    g, g_lagrange = O.params_setup(k, s)
    t_g1, t_lagrange, t_open0 = O.table_srs_setup(max(N, n), s)
    Nt = max(N, n)
    params = cq.ParamsKZG(k, g, g_lagrange)
    table_srs = cq.TableSRS(t_g1, t_lagrange, t_open0)
    dom = cq.EvaluationDomain(3, k)  # required_degree() = 3 for the CQ argument (static_lookup.rs:187-190)
    odom = O.domain_new(3, k)
    assert dom.extended_k == k + 1

    def check_point(got, exp_aff):
        assert np.array_equal(got.to_affine(), exp_aff)
        return O.g1_to_bytes(exp_aff)

    transcript = []
    # ---- two tables of N distinct values; advice rows look up row idx[r] of both tables (vector lookup) ------------
    table1 = [int(v) for v in rng.choice(1 << 20, N, replace=False)]
    table2 = [int(v) + (1 << 21) for v in rng.choice(1 << 20, N, replace=False)]
    rows = [int(v) for v in rng.integers(0, N, usable)]
    adv1 = [table1[r] for r in rows] + [int(v) for v in rng.integers(0, 1 << 60, n - usable)]  # blinded tail rows
    adv2 = [table2[r] for r in rows] + [int(v) for v in rng.integers(0, 1 << 60, n - usable)]
    # plonk/prover.rs:356-360 advice commitments
    for col in (adv1, adv2):
        a = F(col)
        transcript.append(check_point(params.commit_lagrange(a), O.best_multiexp(a, g_lagrange, 2)[1]))
    theta, beta = 0x1234567, 0x7654321  # transcript challenges stand-ins
    # static_lookup/prover.rs:108-126 compress; :165 f_cm; :167-170 m_cm
    f_vals = [(a1 * theta + a2) % P.R_MOD for a1, a2 in zip(adv1, adv2)]
    f = F(f_vals)
    transcript.append(check_point(params.commit_lagrange(f), O.best_multiexp(f, g_lagrange, 2)[1]))
    m = {}
    for r in rows:
        m[r] = m.get(r, 0) + 1
    idx = np.array(sorted(m), dtype=np.uint32)
    mult = F([m[int(i)] for i in idx])
    m_cm = cq.cq.commit_m(table_srs, {int(i): mult[t] for t, i in enumerate(idx)})
    transcript.append(check_point(m_cm, O.sparse_commit(t_lagrange, idx, mult)))
    # :224-257 a_cm / qa_cm / a0_cm over the support of m
    tvals = {int(i): (table1[int(i)] * theta + table2[int(i)]) % P.R_MOD for i in idx}
    a_vals = F([m[int(i)] * pow((tvals[int(i)] + beta) % P.R_MOD, -1, P.R_MOD) % P.R_MOD for i in idx])
    qs = O.synth_bases(0x9595, Nt, 2)  # stand-in for the theta-compressed cached quotient commitments (keygen-time data)
    qs_dev = cq.DeviceBases(qs)
    a_cm, qa_cm, a0_cm = cq.cq.commit_log_derivative_sparse(table_srs, qs_dev, idx, a_vals)
    transcript.append(check_point(a_cm, O.sparse_commit(t_lagrange, idx, a_vals)))
    transcript.append(check_point(qa_cm, O.sparse_commit(qs, idx, a_vals)))
    transcript.append(check_point(a0_cm, O.sparse_commit(t_open0, idx, a_vals)))
    # :261-276 bs = 1/(f_i + beta) on usable rows, 1/beta on the rest; ifft
    beta_inv = pow(beta, -1, P.R_MOD)
    bs = F([pow((fv + beta) % P.R_MOD, -1, P.R_MOD) for fv in f_vals[:usable]] + [beta_inv] * (n - usable))
    b_coeff = dom.lagrange_to_coeff(bs)
    assert np.array_equal(b_coeff, O.lagrange_to_coeff(odom, bs))
    # :279-311 B_0 = (B - B(0))/X ; p_cm over the degree-bound SRS slice (last n-1 powers, my_test.rs:205) ; b0_cm
    b0_bound = t_g1[Nt - (n - 1):]
    b0_bound_dev = cq.DeviceBases(b0_bound)
    b0_cm, p_cm = cq.cq.commit_b0_and_p(params, b0_bound_dev, b_coeff)
    b0 = np.ascontiguousarray(b_coeff[1:])
    transcript.append(check_point(p_cm, O.best_multiexp(b0, b0_bound, 2)[1]))
    b0_full = np.concatenate([b0, np.zeros((1, 4), np.uint64)])
    transcript.append(check_point(b0_cm, O.best_multiexp(b0_full, g, 2)[1]))
    # :326-332 f -> coefficients
    f_coeff = dom.lagrange_to_coeff(f)
    assert np.array_equal(f_coeff, O.lagrange_to_coeff(odom, f))
    # vanishing/prover.rs:58 random polynomial commitment
    rnd = O.synth_scalars(0x4444 + k, n)
    transcript.append(check_point(params.commit(rnd), O.best_multiexp(rnd, g, 2)[1]))
    # plonk/prover.rs:587-603 advice -> coeff ; evaluation.rs:317-334, 535-536 coeff_to_extended (advice x2, CQ x2)
    ext = []
    for col in (F(adv1), F(adv2), f, bs):
        c = dom.lagrange_to_coeff(col)
        oc = O.lagrange_to_coeff(odom, col)
        assert np.array_equal(c, oc)
        e = dom.coeff_to_extended(c)
        assert np.array_equal(e.values, O.coeff_to_extended(odom, oc))
        ext.append(e.values)
    # evaluate_h itself is out of scope (row-wise gate program); any extended-domain vector exercises the quotient path:
    h_ext = O.synth_scalars(0x5555 + k, 1 << dom.extended_k)
    # vanishing/prover.rs:84-107 divide, extended_to_coeff, split into n-sized pieces, commit each
    h_coeff = dom.extended_to_coeff(dom.divide_by_vanishing_poly(cq.domain.ExtendedLagrange(h_ext)))
    oh = O.extended_to_coeff(odom, O.divide_by_vanishing_poly(odom, h_ext))
    assert np.array_equal(h_coeff, oh)
    assert h_coeff.shape[0] == n * dom.quotient_poly_degree
    for piece in range(dom.quotient_poly_degree):
        hp = np.ascontiguousarray(h_coeff[piece * n:(piece + 1) * n])
        transcript.append(check_point(params.commit(hp), O.best_multiexp(hp, g, 2)[1]))
    # GWC witness commitment (gwc/prover.rs:80-86): one more MSM of n-1 coefficients
    wit = O.synth_scalars(0x6666 + k, n - 1)
    transcript.append(check_point(params.commit(wit), O.best_multiexp(wit, g[: n - 1], 2)[1]))
    assert len(transcript) == 13 and all(len(b) == 32 for b in transcript)
    for d in (qs_dev, b0_bound_dev):
        d.free()
    params.free()
    table_srs.free()

"""The commit phase of create_proof (reference plonk/prover.rs:84-584) for a my_test.rs-shaped circuit — two advice columns,
one static (CQ) lookup over two tables, a permutation over the two advice columns — driven by the REAL Fiat-Shamir transcript
(tests/transcript_ref.py: Blake2b, transcript.rs:199-315): advice commitments -> theta -> CQ commit (f, m) -> beta, gamma ->
permutation product commitments -> commit_log_derivatives (A, Q_A, A_0, B_0, P) -> random polynomial -> y.

Two provers run it independently: the device path (libcqb200 through the Python mirror, vectors resident in HBM) and the CPU
oracle's restatement of the reference. With the same SRS, witness and "rng" values they must produce the same proof bytes and
the same challenges at every step — the operational meaning of "byte-identical transcript" (SURVEY.md §8c)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402
from tests.transcript_ref import Blake2bWrite  # noqa: E402


@pytest.fixture(scope="module")
def cq():
    import cqb200

    cqb200._lib.init(0)
    return cqb200


def F(vals):
    return P.fr_array_from_ints(vals)


def L1(x):
    return P.fr_array_from_ints([x])[0]


@pytest.mark.parametrize("k,N", [(6, 64), (9, 128)])
def test_commit_phase_transcripts_agree(cq, oracle, k, N):
    O, L = oracle, cq._lib
    lib = L.lib()
    n, bf, cs_degree = 1 << k, 5, 4
    usable = n - (bf + 1)
    rng = np.random.default_rng(31 + k)
    s = O.synth_scalars(0xF5 + k, 1)[0]
    Nt = max(N, n)
    g, g_lagrange = O.params_setup(k, s)
    t_g1, _, _ = O.table_srs_setup(Nt, s)
    _, t_lag, t_op0 = O.table_srs_setup(N, s)
    tvals = [[int(v) + (j << 40) for v in rng.choice(1 << 30, N, replace=False)] for j in range(2)]
    rows = [int(v) for v in rng.integers(0, N, usable)]
    adv = [[tv[r] for r in rows] + [int(v) for v in rng.integers(0, 1 << 50, n - usable)] for tv in tvals]
    # permutation over the two advice columns: identity except a few swapped equal cells (any sigma gives the same arithmetic)
    omega = P.omega_for(k)
    delta = cq.permutation.FR_DELTA
    sig = [[pow(delta, j, P.R_MOD) * pow(omega, i, P.R_MOD) % P.R_MOD for i in range(n)] for j in range(2)]
    vk_repr = 0x1234ABCD                      # stand-in for pk.vk.transcript_repr (circuit-specific; prover.rs:85)
    rnd_poly = O.synth_scalars(0x4444 + k, n)  # vanishing::Argument::commit's random polynomial (the caller's rng)
    blind_rows = [O.synth_scalars(0xB11D, bf)]
    b0_bound = np.ascontiguousarray(t_g1[Nt - (n - 1):])
    m = {}
    for r in rows:
        m[r] = m.get(r, 0) + 1
    idx = np.array(sorted(m), dtype=np.uint32)
    mult = F([m[int(i)] for i in idx])
    odom = O.domain_new(cs_degree, k)

    # ------------------------------------------------------------------------------------------------- oracle prover
    def oracle_prover():
        t = Blake2bWrite()
        ch = {}
        t.common_scalar(vk_repr)
        for col in adv:                                                       # prover.rs:356-374
            t.write_point(O.best_multiexp(F(col), g_lagrange, 2)[1], O)
        ch["theta"] = theta = t.squeeze_challenge_scalar()                   # :472
        f_int = [(a0 * theta + a1) % P.R_MOD for a0, a1 in zip(*adv)]        # static_lookup/prover.rs:108-121
        t.write_point(O.best_multiexp(F(f_int), g_lagrange, 2)[1], O)         # f_cm  :165, :174
        t.write_point(O.sparse_commit(t_lag, idx, mult), O)                   # m_cm  :167-175
        ch["beta"] = beta = t.squeeze_challenge_scalar()                     # :529
        ch["gamma"] = gamma = t.squeeze_challenge_scalar()                   # :532
        z, _ = O.permutation_product([F(c) for c in adv], [F(sg) for sg in sig], L1(beta), L1(gamma), L1(omega), L1(1), L1(1))
        z[n - bf:] = blind_rows[0]                                            # permutation/prover.rs:152-155
        t.write_point(O.best_multiexp(z, g_lagrange, 2)[1], O)                # :166-176
        qs_host = [O.cq_table_qs(F(v), t_g1[:N] if Nt == N else O.table_srs_setup(N, s)[0], 4) for v in tvals]
        a_acc = qa_acc = a0_acc = None
        add = lambda acc, p: p if acc is None else O.g1_add_jj(acc, p)  # noqa: E731
        for i in idx:                                                         # static_lookup/prover.rs:220-257
            i = int(i)
            values = (tvals[0][i] * theta + tvals[1][i]) % P.R_MOD
            qs = O.g1_to_affine(O.g1_add_ja(O.g1_mul_a(qs_host[0][i], L1(theta)), qs_host[1][i]))
            a_i = L1(m[i] * pow((values + beta) % P.R_MOD, -1, P.R_MOD) % P.R_MOD)
            a_acc, qa_acc, a0_acc = add(a_acc, O.g1_mul_a(t_lag[i], a_i)), add(qa_acc, O.g1_mul_a(qs, a_i)), add(a0_acc, O.g1_mul_a(t_op0[i], a_i))
        beta_inv = pow(beta, -1, P.R_MOD)
        bs = F([pow((fv + beta) % P.R_MOD, -1, P.R_MOD) for fv in f_int[:usable]] + [beta_inv] * (bf + 1))
        b_coeff = O.lagrange_to_coeff(odom, bs)
        b0 = np.ascontiguousarray(b_coeff[1:])
        p_cm = O.best_multiexp(b0, b0_bound, 2)[1]
        b0_cm = O.best_multiexp(np.concatenate([b0, np.zeros((1, 4), np.uint64)]), g, 2)[1]
        for acc in (a_acc, qa_acc, a0_acc):                                    # :301-303
            t.write_point(O.g1_to_affine(acc), O)
        t.write_point(b0_cm, O)                                               # :312
        t.write_point(p_cm, O)                                                # :313
        t.write_point(O.best_multiexp(rnd_poly, g, 2)[1], O)                  # vanishing/prover.rs:58-63
        ch["y"] = t.squeeze_challenge_scalar()                               # prover.rs:584
        return bytes(t.proof), ch

    # ------------------------------------------------------------------------------------------------- device prover
    def device_prover():
        def dev(arr=None, nbytes=None):
            d = ctypes.c_void_p()
            L.check(lib.cqb_dev_alloc(nbytes if arr is None else max(arr.nbytes, 64), ctypes.byref(d)))
            if arr is not None:
                arr = np.ascontiguousarray(arr, dtype=np.uint64)
                L.check(lib.cqb_memcpy_h2d(d, arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes))
            return d

        params = cq.ParamsKZG(k, g, g_lagrange)
        tsrs = cq.TableSRS.setup_from_toxic_waste(N - 1, s, precompute=False)
        big = tsrs if Nt == N else cq.TableSRS.setup_from_toxic_waste(Nt - 1, s, precompute=False)
        tables = [cq.cq.StaticTableValues(F(v), tsrs.g1) for v in tvals]
        bound = cq.DeviceBases(b0_bound)
        d_adv = [dev(F(c)) for c in adv]
        d_sig = [dev(F(sg)) for sg in sig]
        d_f, d_z, d_rnd = dev(nbytes=n * 32), dev(nbytes=n * 32), dev(rnd_poly)
        out, inf = np.zeros(8, np.uint64), ctypes.c_int(0)

        def commit_dev(bases, d_ptr, count):
            L.check(lib.cqb_msm_bn254_g1_dev(bases.handle, 0, d_ptr, count, L.p64(out), ctypes.byref(inf)))
            return out.copy()

        t = Blake2bWrite()
        ch = {}
        t.common_scalar(vk_repr)
        for d in d_adv:
            t.write_point(commit_dev(params.g_lagrange, d, n), O)
        ch["theta"] = theta = t.squeeze_challenge_scalar()
        ptrs = (ctypes.c_void_p * 2)(*d_adv)
        L.check(lib.cqb_fr_compress_dev(ptrs, 2, None, n, L.p64(L1(theta)), d_f))
        t.write_point(commit_dev(params.g_lagrange, d_f, n), O)
        t.write_point(cq.cq.commit_m(tsrs, {int(i): mult[j] for j, i in enumerate(idx)}).to_affine(), O)
        ch["beta"] = beta = t.squeeze_challenge_scalar()
        ch["gamma"] = gamma = t.squeeze_challenge_scalar()
        cq.permutation.commit_dev([d.value for d in d_adv], [d.value for d in d_sig], k, cs_degree, bf, beta, gamma, omega, blind_rows, [d_z.value])
        t.write_point(commit_dev(params.g_lagrange, d_z, n), O)
        cld = cq.cq.commit_log_derivatives_dev(params, tsrs, tables, bound, k, bf, d_f.value, idx, mult, beta, theta)
        for pt in (cld.a_cm, cld.qa_cm, cld.a0_cm, cld.b0_cm, cld.p_cm):
            t.write_point(pt.to_affine(), O)
        t.write_point(commit_dev(params.g, d_rnd, n), O)
        ch["y"] = t.squeeze_challenge_scalar()
        cld.free()
        for d in d_adv + d_sig + [d_f, d_z, d_rnd]:
            L.check(lib.cqb_dev_free(d))
        for tb in tables:
            tb.free()
        bound.free()
        params.free()
        tsrs.free()
        if big is not tsrs:
            big.free()
        return bytes(t.proof), ch

    proof_o, ch_o = oracle_prover()
    proof_d, ch_d = device_prover()
    assert len(proof_o) == 32 * 11                 # 2 advice + f + m + z + A + Q_A + A_0 + B_0 + P + random poly
    assert proof_d == proof_o
    assert ch_d == ch_o and len(set(ch_o.values())) == 4

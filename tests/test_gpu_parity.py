"""GPU parity tests proper: every call goes through the C ABI (libcqb200.so) and is compared bit-for-bit with the CPU
oracle (oracle/bn254_oracle.c, the restatement of the reference's best_multiexp / best_fft / domain / KZG / CQ code) on
the same seeded inputs. Integer work => the bar is exact equality of limbs (NTT) and of the affine normal form (MSM)."""
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


def L(x):
    return P.int_to_limbs(x)


def omega_limbs(k, inverse=False):
    w = P.omega_for(k)
    if inverse:
        w = pow(w, -1, P.R_MOD)
    return L(P.to_mont(w, P.R_MOD))


# ---------------------------------------------------------------------------------------------------------------- NTT
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 13, 16, 17, 18])
def test_best_fft_parity(cq, oracle, log_n):
    n = 1 << log_n
    a = oracle.synth_scalars(0x5EED0002 + log_n, n)
    if n >= 8:  # structured values too: zeros, one, r-1
        a[0] = 0
        a[1] = L(P.MONT % P.R_MOD)
        a[2] = L(P.to_mont(P.R_MOD - 1, P.R_MOD))
    w = omega_limbs(log_n)
    exp = oracle.best_fft(a, w, log_n, threads=4)
    got = a.copy()
    cq.best_fft(got, w, log_n)
    assert np.array_equal(got, exp)


def test_best_fft_wrong_size_panics(cq):
    a = np.zeros((6, 4), np.uint64)
    with pytest.raises(AssertionError):
        cq.best_fft(a, omega_limbs(3), 3)  # arithmetic.rs:184


@pytest.mark.parametrize("j,k", [(3, 3), (3, 4), (4, 5), (5, 6), (3, 10), (4, 12), (6, 9)])
def test_domain_wrappers_parity(cq, oracle, j, k):
    """lagrange_to_coeff, coeff_to_extended, divide_by_vanishing_poly + extended_to_coeff (poly/domain.rs)"""
    od = oracle.domain_new(j, k)
    d = cq.EvaluationDomain(j, k)
    assert d.extended_k == od.extended_k and d.quotient_poly_degree == od.quotient_poly_degree
    for name in ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv", "ifft_divisor",
                 "extended_ifft_divisor"):
        assert np.array_equal(getattr(d, name), od.f(name)), name
    assert np.array_equal(d.t_evaluations, od.t_evals())
    n = 1 << k
    a = oracle.synth_scalars(77 + k, n)
    coeff = d.lagrange_to_coeff(a)
    assert np.array_equal(coeff, oracle.lagrange_to_coeff(od, a))
    ext = d.coeff_to_extended(coeff)
    exp_ext = oracle.coeff_to_extended(od, coeff)
    assert np.array_equal(ext.values, exp_ext)
    # round trip without division
    back = d.extended_to_coeff(ext)
    assert np.array_equal(back, oracle.extended_to_coeff(od, exp_ext))
    assert np.array_equal(back[:n], coeff) and not back[n:].any()
    # quotient path: divide_by_vanishing_poly -> extended_to_coeff (vanishing/prover.rs:84-87)
    h = oracle.synth_scalars(99 + k, 1 << d.extended_k)
    got = d.extended_to_coeff(d.divide_by_vanishing_poly(cq.domain.ExtendedLagrange(h)))
    exp = oracle.extended_to_coeff(od, oracle.divide_by_vanishing_poly(od, h))
    assert np.array_equal(got, exp)


# ---------------------------------------------------------------------------------------------------------------- MSM
def _edge_inputs(oracle, n, seed):
    sc = oracle.synth_scalars(seed, n)
    bases = oracle.synth_bases(seed + 1, n, 4)
    if n >= 12:
        sc[0] = 0                                              # zero scalar
        sc[1] = L(P.to_mont(P.R_MOD - 1, P.R_MOD))             # r - 1
        sc[2] = L(P.to_mont(1, P.R_MOD))
        bases[3] = 0                                           # identity base (0,0)
        bases[5] = bases[4]                                    # repeated point, ...
        sc[5] = sc[4]                                          # ... same scalar => P + P inside a bucket
        bases[7] = oracle.g1_neg_a(bases[6])                   # P + (-P) inside a bucket
        sc[7] = sc[6]
        sc[8] = L(P.to_mont((1 << 253) + 12345, P.R_MOD))      # top window populated
        sc[9] = L(P.to_mont(0xFFFF, P.R_MOD))                  # carries in the signed-digit recoding
        sc[10] = L(P.to_mont(0x8000, P.R_MOD))
        sc[11] = L(P.to_mont((1 << 254) % P.R_MOD, P.R_MOD))
    return sc, bases


@pytest.mark.parametrize("n", [1, 2, 3, 4, 12, 31, 32, 33, 100, 257, 1000, 4096, 5000, 1 << 14, (1 << 16) + 3])
def test_best_multiexp_parity(cq, oracle, n):
    sc, bases = _edge_inputs(oracle, n, 1000 + n)
    _, exp = oracle.best_multiexp(sc, bases, 8)
    got = cq.best_multiexp(sc, bases)
    assert np.array_equal(got.to_affine(), exp)
    assert got.is_identity == (not exp.any())


def test_best_multiexp_empty_and_mismatch(cq):
    got = cq.best_multiexp(np.zeros((0, 4), np.uint64), np.zeros((0, 8), np.uint64))
    assert got.is_identity and not got.to_affine().any()
    with pytest.raises(AssertionError):
        cq.best_multiexp(np.zeros((3, 4), np.uint64), np.zeros((2, 8), np.uint64))  # arithmetic.rs:133


@pytest.mark.parametrize("kind", ["all_zero", "all_equal", "small", "witness_like", "cancel", "negative_small", "few_values", "bits"])
def test_best_multiexp_structured_scalars(cq, oracle, kind):
    n = 3000
    bases = oracle.synth_bases(4242, n, 4)
    sc = oracle.synth_scalars(4243, n)
    rng = np.random.default_rng(5)
    if kind == "all_zero":
        sc[:] = 0
    elif kind == "all_equal":
        sc[:] = sc[0]
    elif kind == "small":
        sc = P.fr_array_from_ints([int(v) for v in rng.integers(0, 1 << 16, n)])
    elif kind == "witness_like":
        vals = [0 if rng.random() < 0.9 else int(rng.integers(0, 4)) for _ in range(n)]
        sc = P.fr_array_from_ints(vals)
    elif kind == "negative_small":  # r - x: identical upper digits in every lane, varying low window
        sc = P.fr_array_from_ints([P.R_MOD - int(v) for v in rng.integers(1, 1 << 10, n)])
    elif kind == "few_values":      # 5 distinct scalars: warps straddle the hot-key threshold of the count/scatter kernels
        sc = sc[rng.integers(0, 5, n)]
    elif kind == "bits":
        sc = P.fr_array_from_ints([int(v) for v in rng.integers(0, 2, n)])
    elif kind == "cancel":  # sum is the identity: s*P + (r-s)*P
        bases[1::2] = bases[0::2]
        ints = P.fr_array_to_ints(sc[0::2])
        sc[1::2] = P.fr_array_from_ints([(P.R_MOD - v) % P.R_MOD for v in ints])
    _, exp = oracle.best_multiexp(sc, bases, 8)
    got = cq.best_multiexp(sc, bases)
    assert np.array_equal(got.to_affine(), exp)
    if kind in ("all_zero", "cancel"):
        assert got.is_identity


@pytest.mark.parametrize("c", [4, 7, 11, 13, 16])
def test_msm_window_bits_do_not_change_result(cq, oracle, c):
    n = 2000
    sc, bases = _edge_inputs(oracle, n, 31337)
    _, exp = oracle.best_multiexp(sc, bases, 8)
    lib = cq._lib.lib()
    cq._lib.check(lib.cqb_msm_set_window_bits(c))
    try:
        got = cq.best_multiexp(sc, bases)
    finally:
        cq._lib.check(lib.cqb_msm_set_window_bits(0))
    assert np.array_equal(got.to_affine(), exp)


def test_kzg_commit_identity_and_parity(cq, oracle):
    """kzg/commitment.rs:570-593 test_commit_lagrange through ParamsKZG on the device, plus parity with the oracle"""
    K = 6
    s = oracle.synth_scalars(0xC0, 1)[0]
    g, gl = oracle.params_setup(K, s)
    params = cq.ParamsKZG(K, g, gl)
    d = cq.EvaluationDomain(1, K)
    a = P.fr_array_from_ints(list(range(1 << K)))
    b = d.lagrange_to_coeff(a)
    c1 = params.commit(b)
    c2 = params.commit_lagrange(a)
    assert c1 == c2
    _, exp = oracle.best_multiexp(a, gl, 4)
    assert np.array_equal(c2.to_affine(), exp)
    short = a[:40]  # size < n is allowed (commitment.rs:502)
    _, exp_s = oracle.best_multiexp(short, gl[:40], 2)
    assert np.array_equal(params.commit_lagrange(short).to_affine(), exp_s)
    with pytest.raises(AssertionError):
        params.commit(np.zeros(((1 << K) + 1, 4), np.uint64))
    params.free()


def test_cq_sparse_commits_parity(cq, oracle):
    """static_lookup/prover.rs:167-170, 245-257: serial scalar-mul loops == one sparse MSM each"""
    N = 64
    s = oracle.synth_scalars(0xC1, 1)[0]
    g1, gl, op0 = oracle.table_srs_setup(N, s)
    srs = cq.TableSRS(g1, gl, op0)
    rng = np.random.default_rng(9)
    idx = np.sort(rng.choice(N, 23, replace=False)).astype(np.uint32)
    mult = P.fr_array_from_ints([int(v) for v in rng.integers(1, 50, idx.shape[0])])
    m_sparse = {int(i): mult[t] for t, i in enumerate(idx)}
    m_cm = cq.cq.commit_m(srs, m_sparse)
    assert np.array_equal(m_cm.to_affine(), oracle.sparse_commit(gl, idx, mult))
    a_vals = oracle.synth_scalars(123, idx.shape[0])
    qs = oracle.synth_bases(555, N, 2)  # stand-in for the cached quotient commitments
    qs_dev = cq.DeviceBases(qs)
    a_cm, qa_cm, a0_cm = cq.cq.commit_log_derivative_sparse(srs, qs_dev, idx, a_vals)
    assert np.array_equal(a_cm.to_affine(), oracle.sparse_commit(gl, idx, a_vals))
    assert np.array_equal(qa_cm.to_affine(), oracle.sparse_commit(qs, idx, a_vals))
    assert np.array_equal(a0_cm.to_affine(), oracle.sparse_commit(op0, idx, a_vals))
    # empty support
    assert cq.cq.commit_m(srs, {}).is_identity
    qs_dev.free()
    srs.free()


def test_synth_generators_match_oracle(cq, oracle):
    lib = cq._lib.lib()
    n = 5000
    d = ctypes.c_void_p()
    cq._lib.check(lib.cqb_dev_alloc(n * 64, ctypes.byref(d)))
    out_s = np.zeros((n, 4), np.uint64)
    cq._lib.check(lib.cqb_synth_scalars_dev(0x5EED0001, 7, n, d))
    cq._lib.check(lib.cqb_memcpy_d2h(out_s.ctypes.data_as(ctypes.c_void_p), d, n * 32))
    assert np.array_equal(out_s, oracle.synth_scalars(0x5EED0001, n, start=7))
    out_b = np.zeros((n, 8), np.uint64)
    cq._lib.check(lib.cqb_synth_bases_dev(0xC0FFEE, 0, n, d))
    cq._lib.check(lib.cqb_memcpy_d2h(out_b.ctypes.data_as(ctypes.c_void_p), d, n * 64))
    assert np.array_equal(out_b, oracle.synth_bases(0xC0FFEE, n, 4))
    cq._lib.check(lib.cqb_dev_free(d))


def test_g1_sum_affine(cq, oracle):
    pts = oracle.synth_bases(77, 9, 1)
    pts[4] = 0
    acc = np.zeros(12, np.uint64)
    for i in range(9):
        acc = oracle.g1_add_ja(acc, pts[i])
    exp = oracle.g1_to_affine(acc)
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    cq._lib.check(cq._lib.lib().cqb_g1_sum_affine(cq._lib.p64(pts), 9, cq._lib.p64(out), ctypes.byref(inf)))
    assert np.array_equal(out, exp) and inf.value == 0


@pytest.mark.parametrize("kind", ["all_equal", "two_values", "top_window_only"])
def test_best_multiexp_long_buckets(cq, oracle, kind):
    """heavily repeated digits: buckets spanning many chunks exercise msm_merge_kernel's long path and msm_merge_big_kernel"""
    n = 1 << 15
    bases = oracle.synth_bases(777, n, 8)
    sc = oracle.synth_scalars(778, n)
    if kind == "all_equal":
        sc[:] = sc[0]
    elif kind == "two_values":
        sc[0::2] = sc[0]
        sc[1::2] = sc[1]
    else:  # only the two top bits of the scalar differ
        base = (1 << 200) + 12345
        sc = P.fr_array_from_ints([((i % 3) << 252) + base for i in range(n)])
    _, exp = oracle.best_multiexp(sc, bases, 8)
    got = cq.best_multiexp(sc, bases)
    assert np.array_equal(got.to_affine(), exp)


@pytest.mark.parametrize("n,c", [(1 << 12, 10), (5000, 12), (1 << 15, 0), (1 << 16, 17), (40000, 20)])
def test_precomputed_table_msm_parity(cq, oracle, n, c):
    """single-bucket-set layout over the per-SRS table 2^(c w) P_i (cqb_bases_precompute): same result as the windowed
    layout and as the oracle, for full-size, prefix, offset and sparse MSMs"""
    sc, bases = _edge_inputs(oracle, n, 9000 + n)
    dev = cq.DeviceBases(bases, precompute=True, window_bits=c)
    _, exp = oracle.best_multiexp(sc, bases, 8)
    assert np.array_equal(dev.msm(sc).to_affine(), exp)
    m = n // 2 + 3                      # prefix (commit of a shorter polynomial, commitment.rs:502)
    _, exp_p = oracle.best_multiexp(sc[:m], bases[:m], 8)
    assert np.array_equal(dev.msm(sc[:m]).to_affine(), exp_p)
    off = n // 3                        # offset slice (degree-bound commit over the tail of the SRS)
    _, exp_o = oracle.best_multiexp(sc[: n - off], bases[off:], 8)
    assert np.array_equal(dev.msm(sc[: n - off], offset=off).to_affine(), exp_o)
    tiny = 17                           # far below 1/8 of the set: falls back to the windowed layout
    _, exp_t = oracle.best_multiexp(sc[:tiny], bases[:tiny], 1)
    assert np.array_equal(dev.msm(sc[:tiny]).to_affine(), exp_t)
    rng = np.random.default_rng(3)
    idx = np.sort(rng.choice(n, n // 2, replace=False)).astype(np.uint32)
    dense = np.zeros((n, 4), np.uint64)
    dense[idx] = sc[: idx.shape[0]]
    _, exp_s = oracle.best_multiexp(dense, bases, 8)
    assert np.array_equal(dev.msm_sparse(idx, sc[: idx.shape[0]]).to_affine(), exp_s)
    # skewed scalars on the single-set layout
    sk = sc.copy()
    sk[:] = sk[5]
    _, exp_k = oracle.best_multiexp(sk, bases, 8)
    assert np.array_equal(dev.msm(sk).to_affine(), exp_k)
    dev.free()


@pytest.mark.parametrize("precompute", [False, True])
def test_pipelined_host_msm_matches_device_path(cq, oracle, precompute):
    """host-pointer MSM from pinned memory is cut into 4 parts (H2D of part p+1 overlaps the kernels of part p): the result
    must equal the device-resident single-shot path and the oracle's result on a sample-sized check of linearity"""
    lib = cq._lib.lib()
    n = (1 << 21) + 12345
    d_b, d_s, h_pin = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    cq._lib.check(lib.cqb_dev_alloc(n * 64, ctypes.byref(d_b)))
    cq._lib.check(lib.cqb_dev_alloc(n * 32, ctypes.byref(d_s)))
    cq._lib.check(lib.cqb_host_alloc_pinned(n * 32, ctypes.byref(h_pin)))
    cq._lib.check(lib.cqb_synth_bases_dev(0xABCD, 0, n, d_b))
    cq._lib.check(lib.cqb_synth_scalars_dev(0xABCE, 0, n, d_s))
    cq._lib.check(lib.cqb_memcpy_d2h(h_pin, d_s, n * 32))
    h = ctypes.c_uint64(0)
    cq._lib.check(lib.cqb_bases_register_device(d_b, n, ctypes.byref(h)))
    if precompute:
        cq._lib.check(lib.cqb_bases_precompute(h.value, 0))
    out_dev, out_pin, out_page = (np.zeros(8, np.uint64) for _ in range(3))
    inf = ctypes.c_int(0)
    cq._lib.check(lib.cqb_msm_bn254_g1_dev(h.value, 0, d_s, n, cq._lib.p64(out_dev), ctypes.byref(inf)))
    cq._lib.check(lib.cqb_msm_bn254_g1(h.value, 0, ctypes.cast(h_pin, cq._lib.u64p), n, cq._lib.p64(out_pin), ctypes.byref(inf)))
    pageable = np.ctypeslib.as_array(ctypes.cast(h_pin, cq._lib.u64p), shape=(n * 4,)).copy()
    cq._lib.check(lib.cqb_msm_bn254_g1(h.value, 0, cq._lib.p64(pageable), n, cq._lib.p64(out_page), ctypes.byref(inf)))
    assert np.array_equal(out_dev, out_pin) and np.array_equal(out_dev, out_page) and out_dev.any()
    # anchor to the oracle: MSM over the first 2^14 points through the same handle
    m = 1 << 14
    sc = pageable.reshape(n, 4)[:m].copy()
    bs = np.zeros((m, 8), np.uint64)
    cq._lib.check(lib.cqb_memcpy_d2h(bs.ctypes.data_as(ctypes.c_void_p), d_b, m * 64))
    _, exp = oracle.best_multiexp(sc, bs, 8)
    out_s = np.zeros(8, np.uint64)
    cq._lib.check(lib.cqb_msm_bn254_g1(h.value, 0, cq._lib.p64(sc), m, cq._lib.p64(out_s), ctypes.byref(inf)))
    assert np.array_equal(out_s, exp)
    cq._lib.check(lib.cqb_bases_free(h.value))
    cq._lib.check(lib.cqb_host_free_pinned(h_pin))
    cq._lib.check(lib.cqb_dev_free(d_b))
    cq._lib.check(lib.cqb_dev_free(d_s))


def test_error_codes_mirror_reference_panics(cq, oracle):
    """the ABI returns CQB_E_* codes where the reference panics; nothing crashes, nothing falls back"""
    lib, L = cq._lib.lib(), cq._lib
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    sc = oracle.synth_scalars(1, 8)
    bs = oracle.synth_bases(2, 8, 1)
    dev = cq.DeviceBases(bs, precompute=False)
    # assert!(self.n() >= size) poly/kzg/commitment.rs:502
    assert lib.cqb_msm_bn254_g1(dev.handle, 0, L.p64(np.zeros((9, 4), np.uint64)), 9, L.p64(out), ctypes.byref(inf)) == L.CQB_E_LEN_MISMATCH
    assert lib.cqb_msm_bn254_g1(dev.handle, 5, L.p64(sc), 4, L.p64(out), ctypes.byref(inf)) == L.CQB_E_LEN_MISMATCH
    assert lib.cqb_msm_bn254_g1(12345678, 0, L.p64(sc), 8, L.p64(out), ctypes.byref(inf)) == L.CQB_E_BAD_ARG
    assert lib.cqb_msm_bn254_g1(dev.handle, 0, None, 8, L.p64(out), ctypes.byref(inf)) == L.CQB_E_BAD_ARG
    bad_idx = np.array([0, 99], np.uint32)  # slice index out of range in the reference's loop
    assert lib.cqb_msm_bn254_g1_sparse(dev.handle, bad_idx.ctypes.data_as(L.u32p), L.p64(sc), 2, L.p64(out), ctypes.byref(inf)) == L.CQB_E_BAD_ARG
    assert lib.cqb_ntt_bn254_fr(L.p64(sc), L.p64(sc[0]), 29) == L.CQB_E_BAD_SIZE  # k <= Fr::S = 28
    assert lib.cqb_bases_precompute(dev.handle, 99) == L.CQB_E_BAD_ARG
    assert b"window bits" in lib.cqb_last_error()
    # still healthy afterwards
    _, exp = oracle.best_multiexp(sc, bs, 1)
    assert np.array_equal(dev.msm(sc).to_affine(), exp)
    dev.free()


def test_concurrent_callers_are_serialised(cq, oracle):
    """nothing forbids concurrent callers (rayon), so the library must be re-entrant: four host threads issue
    MSMs and NTTs at once (ctypes drops the GIL during the calls); every result must match the oracle"""
    import threading

    n, k = 3000, 11
    bs = oracle.synth_bases(50, n, 4)
    dev = cq.DeviceBases(bs, precompute=False)
    w = omega_limbs(k)
    jobs, errors = [], []
    for t in range(4):
        sc = oracle.synth_scalars(60 + t, n)
        a = oracle.synth_scalars(70 + t, 1 << k)
        jobs.append((sc, oracle.best_multiexp(sc, bs, 2)[1], a, oracle.best_fft(a, w, k, 2)))

    def work(job):
        sc, exp_pt, a, exp_fft = job
        try:
            for _ in range(5):
                if not np.array_equal(dev.msm(sc).to_affine(), exp_pt):
                    errors.append("msm")
                b = a.copy()
                cq.best_fft(b, w, k)
                if not np.array_equal(b, exp_fft):
                    errors.append("fft")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    dev.free()


def test_shutdown_and_reinit(cq, oracle):
    lib, L = cq._lib.lib(), cq._lib
    sc, bs = oracle.synth_scalars(80, 100), oracle.synth_bases(81, 100, 1)
    _, exp = oracle.best_multiexp(sc, bs, 1)
    lib.cqb_shutdown()
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    assert lib.cqb_msm_bn254_g1_host(L.p64(bs), L.p64(sc), 100, L.p64(out), ctypes.byref(inf)) == L.CQB_E_NO_DEVICE
    L.check(lib.cqb_init(0))
    assert np.array_equal(cq.best_multiexp(sc, bs).to_affine(), exp)


@pytest.mark.parametrize("n,B,c,pre", [(1 << 12, 5, 10, True), (3000, 3, 0, True), (1 << 15, 8, 17, True), (2000, 4, 0, False)])
def test_batched_msm_matches_individual(cq, oracle, n, B, c, pre):
    """cqb_msm_bn254_g1_batch: B commitments over the same SRS in one pass == B separate best_multiexp calls"""
    bases = oracle.synth_bases(0xBA + n, n, 4)
    dev = cq.DeviceBases(bases, precompute=pre, window_bits=c)
    polys = np.stack([oracle.synth_scalars(0xBB + b, n) for b in range(B)])
    polys[1][:] = polys[1][7]          # a skewed member
    polys[2][:] = 0                    # an all-zero member -> identity
    got = dev.msm_batch(polys)
    for b in range(B):
        _, exp = oracle.best_multiexp(polys[b], bases, 8)
        assert np.array_equal(got[b].to_affine(), exp), b
    assert got[2].is_identity
    m = n - 5                          # shorter polynomials with an offset into the set
    got = dev.msm_batch(np.ascontiguousarray(polys[:, :m]), offset=5)
    for b in range(B):
        _, exp = oracle.best_multiexp(np.ascontiguousarray(polys[b, :m]), bases[5:], 8)
        assert np.array_equal(got[b].to_affine(), exp), b
    dev.free()


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_msm_parts_do_not_change_result(cq, oracle, parts):
    """cqb_msm_set_parts: the point-range parts of a pipelined MSM (own histogram / sorted list / bucket array each, sort on a
    second stream) give the single-shot result, on both layouts and with an offset"""
    lib = cq._lib.lib()
    n = 50021
    sc, bases = _edge_inputs(oracle, n, 4711)
    _, exp = oracle.best_multiexp(sc, bases, 8)
    off = 1234
    _, exp_o = oracle.best_multiexp(sc[: n - off], bases[off:], 8)
    for pre in (False, True):
        dev = cq.DeviceBases(bases, precompute=pre, window_bits=13 if pre else 0)
        cq._lib.check(lib.cqb_msm_set_parts(parts))
        try:
            assert np.array_equal(dev.msm(sc).to_affine(), exp)
            assert np.array_equal(dev.msm(sc[: n - off], offset=off).to_affine(), exp_o)
        finally:
            cq._lib.check(lib.cqb_msm_set_parts(0))
        dev.free()


@pytest.mark.parametrize("n", [0, 1, 2, 17, 100, 1000, 5000])
def test_msmkzg_eval_and_batch_normalize_parity(cq, oracle, n):
    """poly/kzg/msm.rs:65-70 MSMKZG::eval: Jacobian bases -> batch_normalize (derive/curve.rs:362-397) -> best_multiexp"""
    rng = np.random.default_rng(40 + n)
    aff = oracle.synth_bases(0xE7A1 + n, max(n, 1), 2)[:n]
    sc = oracle.synth_scalars(0xE7A2 + n, max(n, 1))[:n]
    # non-trivial z: scale every point by a random scalar in Jacobian form (g1_mul_a returns Jacobian), keep the multiplier's inverse in the scalar
    jac = np.zeros((n, 12), np.uint64)
    for i in range(n):
        jac[i] = oracle.g1_mul_a(aff[i], oracle.synth_scalars(0xE7A3 + i, 1)[0])
    if n >= 17:
        jac[3] = 0                     # identity (z = 0)
        jac[5, 8:] = 0                 # z = 0 with junk x, y: still the identity
        sc[7] = 0
    exp_aff = oracle.g1_batch_normalize(jac) if n else np.zeros((0, 8), np.uint64)
    assert np.array_equal(cq.batch_normalize(jac), exp_aff)
    m = cq.MSMKZG()
    for i in range(n):
        m.append_term(sc[i], jac[i])
    got = m.eval()
    if n:
        _, exp = oracle.best_multiexp(sc, exp_aff, 4)
    else:
        exp = np.zeros(8, np.uint64)
    assert np.array_equal(got.to_affine(), exp)
    # check(): sum s_i P_i + (-sum) == identity
    if n >= 2:
        tot = oracle.g1_mul_a(got.to_affine(), P.int_to_limbs(P.to_mont(P.R_MOD - 1, P.R_MOD))) if got.to_affine().any() else np.zeros(12, np.uint64)
        m.append_term(P.int_to_limbs(P.to_mont(1, P.R_MOD)), tot)
        assert m.check()


def test_generic_best_multiexp_keeps_host_bases_resident(cq, oracle):
    """arithmetic.rs:132 called in a loop over the SAME bases slice (what `commit` does per column): the second call is served by the
    resident copy, the third by its table; a different slice at the same address (or edited points) is a different fingerprint. Results
    never change."""
    lib = cq._lib.lib()
    n = (1 << 16) + 77
    bases = oracle.synth_bases(0xCAC4E, n, 4)
    out0 = lib.cqb_launch_count()
    for rep in range(4):
        sc = oracle.synth_scalars(0xCAC5 + rep, n)
        _, exp = oracle.best_multiexp(sc, bases, 8)
        assert np.array_equal(cq.best_multiexp(sc, bases).to_affine(), exp), rep
    # the same buffer refilled with other points: must not be served from the cache
    bases[:] = oracle.synth_bases(0xCAC4F, n, 4)
    sc = oracle.synth_scalars(0xCAD0, n)
    _, exp = oracle.best_multiexp(sc, bases, 8)
    assert np.array_equal(cq.best_multiexp(sc, bases).to_affine(), exp)
    # cache off: same results
    cq._lib.check(lib.cqb_set_host_bases_cache(0))
    try:
        assert np.array_equal(cq.best_multiexp(sc, bases).to_affine(), exp)
    finally:
        cq._lib.check(lib.cqb_set_host_bases_cache(-1))
    assert lib.cqb_launch_count() > out0

"""BASELINE.json full sizes (2^24), where the CPU oracle would take minutes: size-independent properties the domain offers.
  MSM : doubling every scalar doubles the result (linearity); splitting the point range and folding the partials gives the
        same point (the multi-GPU decomposition, arithmetic.rs:137-153); windowed and table layouts agree; a 2^20 prefix is
        anchored to the oracle.
  NTT : inverse(forward(a)) == a; sum_k NTT(a)[k] = n * a[0]; coset round trip with the vanishing division undone.
Everything stays on the device; only 64-byte points and a few field elements come back."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402

LOG_N = 24


@pytest.fixture(scope="module")
def env(oracle):
    import cqb200

    cqb200._lib.init(0)
    L, lib = cqb200._lib, cqb200._lib.lib()
    n = 1 << LOG_N

    def dalloc(nbytes):
        d = ctypes.c_void_p()
        L.check(lib.cqb_dev_alloc(nbytes, ctypes.byref(d)))
        return d

    d_b, d_s, d_s2 = dalloc(n * 64), dalloc(n * 32), dalloc(n * 32)
    L.check(lib.cqb_synth_bases_dev(0xC0FFEE, 0, n, d_b))
    L.check(lib.cqb_synth_scalars_dev(0x5EED0001, 0, n, d_s))
    yield cqb200, L, lib, n, d_b, d_s, d_s2
    for d in (d_b, d_s, d_s2):
        L.check(lib.cqb_dev_free(d))


def _msm(L, lib, h, d_s, n, offset=0):
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    L.check(lib.cqb_msm_bn254_g1_dev(h, offset, d_s, n, L.p64(out), ctypes.byref(inf)))
    return out


def test_msm_2p24_properties(env, oracle):
    cq, L, lib, n, d_b, d_s, d_s2 = env
    h = ctypes.c_uint64(0)
    L.check(lib.cqb_bases_register_device(d_b, n, ctypes.byref(h)))
    r_windowed = _msm(L, lib, h.value, d_s, n)
    assert r_windowed.any()
    # range split + fold == whole (what sharded.py does across GPUs)
    parts = np.stack([_msm(L, lib, h.value, ctypes.c_void_p(d_s.value + i * (n // 4) * 32), n // 4, offset=i * (n // 4)) for i in range(4)])
    fold = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    L.check(lib.cqb_g1_sum_affine(L.p64(parts), 4, L.p64(fold), ctypes.byref(inf)))
    assert np.array_equal(fold, r_windowed)
    # table layout agrees with the windowed layout
    L.check(lib.cqb_bases_precompute(h.value, 0))
    r_table = _msm(L, lib, h.value, d_s, n)
    assert np.array_equal(r_table, r_windowed)
    # linearity: MSM(2 s, P) == MSM(s, P) + MSM(s, P)
    two = P.int_to_limbs(P.to_mont(2, P.R_MOD))
    L.check(lib.cqb_memcpy_d2d(d_s2, d_s, n * 32))
    L.check(lib.cqb_fr_scale_dev(d_s2, n, L.p64(two)))
    dbl = np.stack([r_table, r_table])
    exp = np.zeros(8, np.uint64)
    L.check(lib.cqb_g1_sum_affine(L.p64(dbl), 2, L.p64(exp), ctypes.byref(inf)))
    assert np.array_equal(_msm(L, lib, h.value, d_s2, n), exp)
    # anchor: the first 2^18 points against the CPU oracle
    m = 1 << 18
    sc = np.zeros((m, 4), np.uint64)
    bs = np.zeros((m, 8), np.uint64)
    L.check(lib.cqb_memcpy_d2h(sc.ctypes.data_as(ctypes.c_void_p), d_s, m * 32))
    L.check(lib.cqb_memcpy_d2h(bs.ctypes.data_as(ctypes.c_void_p), d_b, m * 64))
    _, cpu = oracle.best_multiexp(sc, bs, oracle.hw_threads())
    assert np.array_equal(_msm(L, lib, h.value, d_s, m), cpu)
    L.check(lib.cqb_bases_free(h.value))


def test_ntt_2p24_properties(env, oracle):
    cq, L, lib, n, d_b, d_s, d_s2 = env
    k = LOG_N
    d = cq.EvaluationDomain(3, k - 1)  # extended domain of size 2^24
    dom = cq.EvaluationDomain(1, k)
    a0 = np.zeros((4, 4), np.uint64)
    L.check(lib.cqb_memcpy_d2h(a0.ctypes.data_as(ctypes.c_void_p), d_s, 128))
    L.check(lib.cqb_memcpy_d2d(d_s2, d_s, n * 32))
    L.check(lib.cqb_ntt_bn254_fr_dev(d_s2, L.p64(dom.omega), k))
    # sum_k A[k] = n * a[0]: evaluate the "polynomial" A at 1
    one = P.int_to_limbs(P.MONT % P.R_MOD)
    ssum = np.zeros(4, np.uint64)
    L.check(lib.cqb_eval_polynomial_dev(d_s2, n, L.p64(one), L.p64(ssum)))
    n_fr = P.int_to_limbs(P.to_mont(n, P.R_MOD))
    assert np.array_equal(ssum, oracle.fr_op("mul", n_fr, a0[0]))
    # A[0] = sum_j a[j]
    asum = np.zeros(4, np.uint64)
    L.check(lib.cqb_eval_polynomial_dev(d_s, n, L.p64(one), L.p64(asum)))
    first = np.zeros(4, np.uint64)
    L.check(lib.cqb_memcpy_d2h(first.ctypes.data_as(ctypes.c_void_p), d_s2, 32))
    assert np.array_equal(first, asum)
    # inverse(forward(a)) == a : compare through a random evaluation point (Schwartz-Zippel) and the first elements
    L.check(lib.cqb_intt_bn254_fr_dev(d_s2, L.p64(dom.omega_inv), L.p64(dom.ifft_divisor), k))
    x = oracle.synth_scalars(0x77, 1)[0]
    e1, e2 = np.zeros(4, np.uint64), np.zeros(4, np.uint64)
    L.check(lib.cqb_eval_polynomial_dev(d_s, n, L.p64(x), L.p64(e1)))
    L.check(lib.cqb_eval_polynomial_dev(d_s2, n, L.p64(x), L.p64(e2)))
    assert np.array_equal(e1, e2)
    back = np.zeros((4, 4), np.uint64)
    L.check(lib.cqb_memcpy_d2h(back.ctypes.data_as(ctypes.c_void_p), d_s2, 128))
    assert np.array_equal(back, a0)
    # coset round trip n/2 coefficients -> 2^24 coset evaluations -> back (no vanishing division): first half == input, rest zero
    half = n // 2
    L.check(lib.cqb_coset_ntt_bn254_fr_dev(d_s, half, d_s2, L.p64(d.extended_omega), d.extended_k, L.p64(d.g_coset), L.p64(d.g_coset_inv)))
    L.check(lib.cqb_coset_intt_bn254_fr_dev(d_s2, d.extended_k, L.p64(d.extended_omega_inv), L.p64(d.extended_ifft_divisor), L.p64(d.g_coset),
                                            L.p64(d.g_coset_inv), None, 0))
    L.check(lib.cqb_eval_polynomial_dev(d_s, half, L.p64(x), L.p64(e1)))
    L.check(lib.cqb_eval_polynomial_dev(d_s2, n, L.p64(x), L.p64(e2)))
    assert np.array_equal(e1, e2)
    tail = np.ones((4, 4), np.uint64)
    L.check(lib.cqb_memcpy_d2h(tail.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(d_s2.value + (n - 4) * 32), 128))
    assert not tail.any()


def test_msm_and_ntt_2p26_properties(oracle):
    """the top of BASELINE.json's sweeps (2^26): table-layout MSM == fold of four range parts, linearity; NTT round trip"""
    import cqb200

    cqb200._lib.init(0)
    L, lib = cqb200._lib, cqb200._lib.lib()
    k = 26
    n = 1 << k

    def dalloc(nbytes):
        d = ctypes.c_void_p()
        L.check(lib.cqb_dev_alloc(nbytes, ctypes.byref(d)))
        return d

    d_b, d_s, d_s2 = dalloc(n * 64), dalloc(n * 32), dalloc(n * 32)
    try:
        L.check(lib.cqb_synth_bases_dev(0xC0FFEE, 0, n, d_b))
        L.check(lib.cqb_synth_scalars_dev(0x5EED0001, 0, n, d_s))
        h = ctypes.c_uint64(0)
        L.check(lib.cqb_bases_register_device(d_b, n, ctypes.byref(h)))
        L.check(lib.cqb_bases_precompute(h.value, 0))
        whole = _msm(L, lib, h.value, d_s, n)
        parts = np.stack([_msm(L, lib, h.value, ctypes.c_void_p(d_s.value + i * (n // 4) * 32), n // 4, offset=i * (n // 4)) for i in range(4)])
        fold = np.zeros(8, np.uint64)
        inf = ctypes.c_int(0)
        L.check(lib.cqb_g1_sum_affine(L.p64(parts), 4, L.p64(fold), ctypes.byref(inf)))
        assert whole.any() and np.array_equal(fold, whole)
        three = P.int_to_limbs(P.to_mont(3, P.R_MOD))
        L.check(lib.cqb_memcpy_d2d(d_s2, d_s, n * 32))
        L.check(lib.cqb_fr_scale_dev(d_s2, n, L.p64(three)))
        exp = np.zeros(8, np.uint64)
        L.check(lib.cqb_g1_sum_affine(L.p64(np.stack([whole, whole, whole])), 3, L.p64(exp), ctypes.byref(inf)))
        assert np.array_equal(_msm(L, lib, h.value, d_s2, n), exp)
        L.check(lib.cqb_bases_free(h.value))
        # NTT 2^26: inverse(forward(a)) == a through a random evaluation point, and A[0] = sum a
        dom = cqb200.EvaluationDomain(1, k)
        x = oracle.synth_scalars(0x78, 1)[0]
        one = P.int_to_limbs(P.MONT % P.R_MOD)
        e1, e2, asum, first = (np.zeros(4, np.uint64) for _ in range(4))
        L.check(lib.cqb_memcpy_d2d(d_s2, d_s, n * 32))
        L.check(lib.cqb_ntt_bn254_fr_dev(d_s2, L.p64(dom.omega), k))
        L.check(lib.cqb_eval_polynomial_dev(d_s, n, L.p64(one), L.p64(asum)))
        L.check(lib.cqb_memcpy_d2h(first.ctypes.data_as(ctypes.c_void_p), d_s2, 32))
        assert np.array_equal(first, asum)
        L.check(lib.cqb_intt_bn254_fr_dev(d_s2, L.p64(dom.omega_inv), L.p64(dom.ifft_divisor), k))
        L.check(lib.cqb_eval_polynomial_dev(d_s, n, L.p64(x), L.p64(e1)))
        L.check(lib.cqb_eval_polynomial_dev(d_s2, n, L.p64(x), L.p64(e2)))
        assert np.array_equal(e1, e2)
    finally:
        for d in (d_b, d_s, d_s2):
            L.check(lib.cqb_dev_free(d))

"""The batched-affine bucket accumulation (msm_accumulate_affine_kernel: the reference's batch_add,
arithmetic/curves/src/derive/curve.rs:4-141, as a B200 kernel) forced on at sizes the oracle finishes in seconds — the
automatic choice only takes it for large MSMs (tests/test_gpu_bigsize_oracle.py covers those). Every exceptional branch of the
affine addition is driven on purpose: identity bases, empty accumulators, P + P (the doubling joins the shared inversion),
P + (-P), buckets that end inside a stream, buckets that span many streams (the long-bucket merge), every stream length."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402


@pytest.fixture(scope="module")
def cq():
    import cqb200

    cqb200._lib.init(0)
    yield cqb200
    cqb200._lib.check(cqb200._lib.lib().cqb_msm_set_accumulator(0, 0))


@pytest.fixture(params=[3, 5, 7], ids=lambda s: f"seg{1 << s}")
def affine(cq, request):
    lib = cq._lib.lib()
    cq._lib.check(lib.cqb_msm_set_accumulator(2, request.param))
    yield request.param
    cq._lib.check(lib.cqb_msm_set_accumulator(0, 0))


def L(x):
    return P.int_to_limbs(x)


def _edge_inputs(oracle, n, seed):
    sc = oracle.synth_scalars(seed, n)
    bases = oracle.synth_bases(seed + 1, n, 4)
    if n >= 12:
        sc[0] = 0
        sc[1] = L(P.to_mont(P.R_MOD - 1, P.R_MOD))
        sc[2] = L(P.to_mont(1, P.R_MOD))
        bases[3] = 0                                           # identity base
        bases[5] = bases[4]                                    # P + P inside a bucket
        sc[5] = sc[4]
        bases[7] = oracle.g1_neg_a(bases[6])                   # P + (-P) inside a bucket
        sc[7] = sc[6]
        sc[8] = L(P.to_mont((1 << 253) + 12345, P.R_MOD))
        sc[9] = L(P.to_mont(0xFFFF, P.R_MOD))
        sc[10] = L(P.to_mont(0x8000, P.R_MOD))
        sc[11] = L(P.to_mont((1 << 254) % P.R_MOD, P.R_MOD))
    return sc, bases


@pytest.mark.parametrize("n", [1, 2, 3, 12, 33, 257, 1000, 5000, (1 << 14) + 7])
def test_affine_windowed_parity(cq, oracle, affine, n):
    sc, bases = _edge_inputs(oracle, n, 2000 + n)
    _, exp = oracle.best_multiexp(sc, bases, 8)
    got = cq.best_multiexp(sc, bases)
    assert np.array_equal(got.to_affine(), exp)


@pytest.mark.parametrize("kind", ["all_zero", "all_equal", "small", "witness_like", "cancel", "negative_small", "few_values", "bits",
                                  "same_point", "pairs_cancel_in_bucket", "identity_bases"])
def test_affine_structured(cq, oracle, affine, kind):
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
        sc = P.fr_array_from_ints([0 if rng.random() < 0.9 else int(rng.integers(0, 4)) for _ in range(n)])
    elif kind == "negative_small":
        sc = P.fr_array_from_ints([P.R_MOD - int(v) for v in rng.integers(1, 1 << 10, n)])
    elif kind == "few_values":
        sc = sc[rng.integers(0, 5, n)]
    elif kind == "bits":
        sc = P.fr_array_from_ints([int(v) for v in rng.integers(0, 2, n)])
    elif kind == "cancel":
        bases[1::2] = bases[0::2]
        ints = P.fr_array_to_ints(sc[0::2])
        sc[1::2] = P.fr_array_from_ints([(P.R_MOD - v) % P.R_MOD for v in ints])
    elif kind == "same_point":          # one point, one scalar: every addition in every bucket is a doubling or hits 2^k P + 2^k P
        bases[:] = bases[0]
        sc[:] = sc[0]
    elif kind == "pairs_cancel_in_bucket":  # P, -P adjacent with the same scalar: accumulators empty out again and again
        bases[1::2] = np.stack([oracle.g1_neg_a(b) for b in bases[0::2]])
        sc[1::2] = sc[0::2]
        sc[:] = sc[rng.integers(0, 3, n) * 2]  # few distinct scalars -> long buckets of cancelling pairs
    elif kind == "identity_bases":
        bases[rng.random(n) < 0.5] = 0
    _, exp = oracle.best_multiexp(sc, bases, 8)
    got = cq.best_multiexp(sc, bases)
    assert np.array_equal(got.to_affine(), exp)


@pytest.mark.parametrize("kind", ["all_equal", "two_values", "top_window_only"])
def test_affine_long_buckets(cq, oracle, affine, kind):
    n = 1 << 15
    bases = oracle.synth_bases(777, n, 8)
    sc = oracle.synth_scalars(778, n)
    if kind == "all_equal":
        sc[:] = sc[0]
    elif kind == "two_values":
        sc[0::2] = sc[0]
        sc[1::2] = sc[1]
    else:
        base = (1 << 200) + 12345
        sc = P.fr_array_from_ints([((i % 3) << 252) + base for i in range(n)])
    _, exp = oracle.best_multiexp(sc, bases, 8)
    got = cq.best_multiexp(sc, bases)
    assert np.array_equal(got.to_affine(), exp)


@pytest.mark.parametrize("n,c", [(1 << 12, 10), (5000, 12), (1 << 16, 17), (40000, 20)])
def test_affine_table_layout_parity(cq, oracle, affine, n, c):
    """single bucket set over the precomputed table, prefix / offset / sparse / batched MSMs"""
    sc, bases = _edge_inputs(oracle, n, 9100 + n)
    dev = cq.DeviceBases(bases, precompute=True, window_bits=c)
    try:
        _, exp = oracle.best_multiexp(sc, bases, 8)
        assert np.array_equal(dev.msm(sc).to_affine(), exp)
        m = n // 2 + 3
        _, exp_p = oracle.best_multiexp(sc[:m], bases[:m], 8)
        assert np.array_equal(dev.msm(sc[:m]).to_affine(), exp_p)
        off = n // 3
        _, exp_o = oracle.best_multiexp(sc[: n - off], bases[off:], 8)
        assert np.array_equal(dev.msm(sc[: n - off], offset=off).to_affine(), exp_o)
        rng = np.random.default_rng(3)
        idx = np.sort(rng.choice(n, n // 2, replace=False)).astype(np.uint32)
        dense = np.zeros((n, 4), np.uint64)
        dense[idx] = sc[: idx.shape[0]]
        _, exp_s = oracle.best_multiexp(dense, bases, 8)
        assert np.array_equal(dev.msm_sparse(idx, sc[: idx.shape[0]]).to_affine(), exp_s)
        sk = sc.copy()
        sk[:] = sk[5]
        _, exp_k = oracle.best_multiexp(sk, bases, 8)
        assert np.array_equal(dev.msm(sk).to_affine(), exp_k)
        # batched MSMs share one launch sequence: one bucket set per member
        lib = cq._lib.lib()
        B = 3
        scs = np.stack([oracle.synth_scalars(500 + b, n) for b in range(B)])
        out = np.zeros((B, 8), np.uint64)
        infs = (ctypes.c_int * B)()
        cq._lib.check(lib.cqb_msm_bn254_g1_batch(dev.handle, 0, cq._lib.p64(scs), n, B, cq._lib.p64(out), infs))
        for b in range(B):
            _, e = oracle.best_multiexp(scs[b], bases, 8)
            assert np.array_equal(out[b], e)
    finally:
        dev.free()


def test_affine_equals_xyzz_at_2p22(cq, oracle):
    """device-resident 2^22 MSM: both accumulation variants, both layouts, one point"""
    L_, lib = cq._lib, cq._lib.lib()
    n = 1 << 22
    d_b, d_s = ctypes.c_void_p(), ctypes.c_void_p()
    L_.check(lib.cqb_dev_alloc(n * 64, ctypes.byref(d_b)))
    L_.check(lib.cqb_dev_alloc(n * 32, ctypes.byref(d_s)))
    try:
        L_.check(lib.cqb_synth_bases_dev(0xC0FFEE, 0, n, d_b))
        L_.check(lib.cqb_synth_scalars_dev(0x5EED0001, 0, n, d_s))
        h = ctypes.c_uint64(0)
        L_.check(lib.cqb_bases_register_device(d_b, n, ctypes.byref(h)))
        res = {}
        for layout in ("windowed", "table"):
            if layout == "table":
                L_.check(lib.cqb_bases_precompute(h.value, 0))
            for mode in (1, 2):
                L_.check(lib.cqb_msm_set_accumulator(mode, 0))
                out = np.zeros(8, np.uint64)
                inf = ctypes.c_int(0)
                L_.check(lib.cqb_msm_bn254_g1_dev(h.value, 0, d_s, n, L_.p64(out), ctypes.byref(inf)))
                res[(layout, mode)] = out
        L_.check(lib.cqb_msm_set_accumulator(0, 0))
        L_.check(lib.cqb_bases_free(h.value))
        first = res[("windowed", 1)]
        assert first.any()
        for k, v in res.items():
            assert np.array_equal(v, first), k
    finally:
        L_.check(lib.cqb_msm_set_accumulator(0, 0))
        L_.check(lib.cqb_dev_free(d_b))
        L_.check(lib.cqb_dev_free(d_s))

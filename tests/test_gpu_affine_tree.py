"""The affine-tree bucket accumulation (msm_affine_tree_kernel + the padded bucket-sorted list; the CPU form is the reference's
batch_add, arithmetic/curves/src/derive/curve.rs:4-141) forced on over the precomputed-table layout at sizes the oracle finishes in
seconds, every tree depth. The exceptional branches are driven on purpose: sentinel padding, identity bases, P + P (joins the shared
inversion as a doubling) at every level (equal points with equal scalars meet again as 2P + 2P, 4P + 4P ...), P + (-P) (identity
results that travel up the tree), empty buckets, runs shorter than one padded unit, prefixes / offsets / sparse index lists, and the
pipelined host-pointer call (several parts, each with its own padded list)."""
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


@pytest.fixture(params=[1, 2, 3, 4, 5], ids=lambda l: f"levels{l}")
def tree(cq, request):
    lib = cq._lib.lib()
    cq._lib.check(lib.cqb_msm_set_tree_levels(request.param))
    cq._lib.check(lib.cqb_msm_set_accumulator(3, 0))
    yield request.param
    cq._lib.check(lib.cqb_msm_set_accumulator(0, 0))
    cq._lib.check(lib.cqb_msm_set_tree_levels(4))


def L(x):
    return P.int_to_limbs(x)


def _edge_inputs(oracle, n, seed):
    sc = oracle.synth_scalars(seed, n)
    bases = oracle.synth_bases(seed + 1, n, 4)
    if n >= 12:
        sc[0] = 0
        sc[1] = L(P.to_mont(P.R_MOD - 1, P.R_MOD))
        sc[2] = L(P.to_mont(1, P.R_MOD))
        bases[3] = 0
        bases[5] = bases[4]
        sc[5] = sc[4]
        bases[7] = oracle.g1_neg_a(bases[6])
        sc[7] = sc[6]
        sc[8] = L(P.to_mont((1 << 253) + 12345, P.R_MOD))
        sc[9] = L(P.to_mont(0xFFFF, P.R_MOD))
        sc[10] = L(P.to_mont(0x8000, P.R_MOD))
        sc[11] = L(P.to_mont((1 << 254) % P.R_MOD, P.R_MOD))
    return sc, bases


@pytest.mark.parametrize("n,c", [(700, 8), (3000, 8), (5000, 9), ((1 << 14) + 7, 8), (1 << 16, 11)])
def test_tree_table_layout_parity(cq, oracle, tree, n, c):
    sc, bases = _edge_inputs(oracle, n, 9300 + n)
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
    finally:
        dev.free()


@pytest.mark.parametrize("kind", ["all_zero", "all_equal", "small", "witness_like", "cancel", "negative_small", "few_values", "bits",
                                  "same_point", "pairs_cancel_in_bucket", "identity_bases", "two_values"])
def test_tree_structured(cq, oracle, tree, kind):
    n = 4000
    bases = oracle.synth_bases(4342, n, 4)
    sc = oracle.synth_scalars(4343, n)
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
    elif kind == "same_point":          # one point, one scalar: P + P, then 2P + 2P, 4P + 4P ... up the tree
        bases[:] = bases[0]
        sc[:] = sc[0]
    elif kind == "pairs_cancel_in_bucket":  # P, -P adjacent with the same scalar: identities travel up the tree
        bases[1::2] = np.stack([oracle.g1_neg_a(b) for b in bases[0::2]])
        sc[1::2] = sc[0::2]
        sc[:] = sc[rng.integers(0, 3, n) * 2]
    elif kind == "identity_bases":
        bases[rng.random(n) < 0.5] = 0
    elif kind == "two_values":
        sc[0::2] = sc[0]
        sc[1::2] = sc[1]
    dev = cq.DeviceBases(bases, precompute=True, window_bits=8)
    try:
        _, exp = oracle.best_multiexp(sc, bases, 8)
        assert np.array_equal(dev.msm(sc).to_affine(), exp)
    finally:
        dev.free()


def test_tree_host_pointer_parts(cq, oracle, tree):
    """cqb_msm_bn254_g1 with host scalars: the call is cut into parts, each sorted into its own padded list"""
    n = 1 << 15
    sc, bases = _edge_inputs(oracle, n, 9500)
    dev = cq.DeviceBases(bases, precompute=True, window_bits=9)
    lib = cq._lib.lib()
    try:
        _, exp = oracle.best_multiexp(sc, bases, 8)
        for parts in (1, 2, 3, 5):
            cq._lib.check(lib.cqb_msm_set_parts(parts))
            assert np.array_equal(dev.msm(sc).to_affine(), exp)
    finally:
        cq._lib.check(lib.cqb_msm_set_parts(0))
        dev.free()


@pytest.mark.parametrize("kind", ["all_equal", "few_values", "small", "negative_small"])
def test_tree_automatic_choice_on_skewed_scalars(cq, oracle, kind):
    """2^21 points: the automatic choice (no forcing) takes the tree from 24 M list entries on, whatever the digits look like — a handful of
    giant buckets (the long-bucket merge behind the tree), or most windows empty"""
    lib = cq._lib.lib()
    cq._lib.check(lib.cqb_msm_set_accumulator(0, 0))
    n = 1 << 21
    bases = oracle.synth_bases(7001, n, 8)
    sc = oracle.synth_scalars(7002, n)
    rng = np.random.default_rng(11)
    if kind == "all_equal":
        sc[:] = sc[0]
    elif kind == "few_values":
        sc = sc[rng.integers(0, 7, n)]
    elif kind == "small":                                    # 16-bit values: every window but the lowest is empty
        small = P.fr_array_from_ints(list(range(1 << 16)))
        sc = small[rng.integers(0, 1 << 16, n)]
    elif kind == "negative_small":
        sc[:] = oracle.synth_scalars(7003, 1)[0]
        sc[::2] = L(P.to_mont(P.R_MOD - 5, P.R_MOD))
    import ctypes

    Lb = cq._lib
    dev = cq.DeviceBases(bases, precompute=True)
    d_s = ctypes.c_void_p()
    Lb.check(lib.cqb_dev_alloc(n * 32, ctypes.byref(d_s)))
    try:
        _, exp = oracle.best_multiexp(sc, bases, oracle.hw_threads())
        sc = np.ascontiguousarray(sc)
        Lb.check(lib.cqb_memcpy_h2d(d_s, sc.ctypes.data_as(ctypes.c_void_p), n * 32))
        out, inf = np.zeros(8, np.uint64), ctypes.c_int(0)
        Lb.check(lib.cqb_msm_bn254_g1_dev(dev.handle, 0, d_s, n, Lb.p64(out), ctypes.byref(inf)))  # device-resident scalars: one part
        assert lib.cqb_msm_last_tree_levels() >= 2
        assert np.array_equal(out, exp)
    finally:
        Lb.check(lib.cqb_dev_free(d_s))
        dev.free()

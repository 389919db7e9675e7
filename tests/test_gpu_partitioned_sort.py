"""The partitioned sort of the MSM digits (msm_part_kernel / msm_part_bin_kernel: tile-local sort by the top 9 bits of the bucket id, then
per-bin count and placement) forced on at sizes the oracle finishes in seconds — the automatic choice only takes it from 4 M list entries
on (tests/test_gpu_bigsize_oracle.py covers those). Window sizes 11..16 (2 to 64 fine ids per coarse bin, both scalars-per-thread
settings), ragged tile ends, hot buckets (warp-aggregated shared-memory atomics), empty bins, prefixes / offsets / sparse index lists,
host-pointer parts, and both accumulations behind it (XYZZ and the affine tree with its padded bucket runs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402


@pytest.fixture(scope="module")
def cq():
    import cqb200

    cqb200._lib.init(0)
    yield cqb200
    lib = cqb200._lib.lib()
    cqb200._lib.check(lib.cqb_msm_set_sort_mode(0))
    cqb200._lib.check(lib.cqb_msm_set_accumulator(0, 0))


@pytest.fixture(params=["xyzz", "tree2", "tree4"])
def part(cq, request):
    lib = cq._lib.lib()
    cq._lib.check(lib.cqb_msm_set_sort_mode(2))
    if request.param == "xyzz":
        cq._lib.check(lib.cqb_msm_set_accumulator(1, 0))
    else:
        cq._lib.check(lib.cqb_msm_set_tree_levels(int(request.param[4:])))
        cq._lib.check(lib.cqb_msm_set_accumulator(3, 0))
    yield request.param
    cq._lib.check(lib.cqb_msm_set_sort_mode(0))
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


@pytest.mark.parametrize("n,c", [(1, 11), (300, 11), (767, 12), (769, 13), (5000, 11), ((1 << 14) + 7, 14), (1 << 16, 16), (40000, 17), (30000, 20)])
def test_partitioned_sort_parity(cq, oracle, part, n, c):
    sc, bases = _edge_inputs(oracle, n, 9700 + n)
    dev = cq.DeviceBases(bases, precompute=True, window_bits=c)
    try:
        _, exp = oracle.best_multiexp(sc, bases, 8)
        assert np.array_equal(dev.msm(sc).to_affine(), exp)
        if n >= 300:
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


@pytest.mark.parametrize("kind", ["all_zero", "all_equal", "small", "witness_like", "negative_small", "few_values", "bits", "same_point",
                                  "two_values", "top_window_only"])
def test_partitioned_sort_structured(cq, oracle, part, kind):
    n = 6000
    bases = oracle.synth_bases(4442, n, 4)
    sc = oracle.synth_scalars(4443, n)
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
    elif kind == "same_point":
        bases[:] = bases[0]
        sc[:] = sc[0]
    elif kind == "two_values":
        sc[0::2] = sc[0]
        sc[1::2] = sc[1]
    elif kind == "top_window_only":
        base = (1 << 200) + 12345
        sc = P.fr_array_from_ints([((i % 3) << 252) + base for i in range(n)])
    dev = cq.DeviceBases(bases, precompute=True, window_bits=12)
    try:
        _, exp = oracle.best_multiexp(sc, bases, 8)
        assert np.array_equal(dev.msm(sc).to_affine(), exp)
    finally:
        dev.free()


def test_partitioned_sort_host_pointer_parts(cq, oracle, part):
    n = 1 << 15
    sc, bases = _edge_inputs(oracle, n, 9900)
    dev = cq.DeviceBases(bases, precompute=True, window_bits=13)
    lib = cq._lib.lib()
    try:
        _, exp = oracle.best_multiexp(sc, bases, 8)
        for parts in (1, 2, 3, 5):
            cq._lib.check(lib.cqb_msm_set_parts(parts))
            assert np.array_equal(dev.msm(sc).to_affine(), exp)
    finally:
        cq._lib.check(lib.cqb_msm_set_parts(0))
        dev.free()

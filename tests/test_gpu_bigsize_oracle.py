"""BASELINE.json sizes compared LIMB FOR LIMB with the CPU oracle (no property shortcuts): the NTT family at 2^20 ... 2^26
(including the 4-pass schedule that starts at log_n = 25 and the coset / coset-inverse wrappers above k = 12) and the MSM at
the full 2^24 (both layouts, the pinned 3-part host-pointer path, the pageable host-pointer path) and 2^26.
Reference: halo2_proofs/src/arithmetic.rs:132-159 (best_multiexp), :171-234 (best_fft), poly/domain.rs:238-338.
The oracle runs on the box's host cores: the 2^24 MSM is ~12 s, a 2^26 transform ~25 s."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402


@pytest.fixture(scope="module")
def env(oracle):
    import cqb200

    cqb200._lib.init(0)
    return cqb200, cqb200._lib, cqb200._lib.lib()


class Dev:
    """device buffers freed at the end of a test"""

    def __init__(self, L, lib):
        self.L, self.lib, self.ptrs = L, lib, []

    def alloc(self, nbytes):
        d = ctypes.c_void_p()
        self.L.check(self.lib.cqb_dev_alloc(max(nbytes, 64), ctypes.byref(d)))
        self.ptrs.append(d)
        return d

    def free(self):
        for d in self.ptrs:
            self.L.check(self.lib.cqb_dev_free(d))
        self.ptrs = []


def _d2h(L, lib, d, shape):
    a = np.empty(shape, np.uint64)
    L.check(lib.cqb_memcpy_d2h(a.ctypes.data_as(ctypes.c_void_p), d, a.nbytes))
    return a


def _h2d(L, lib, d, a):
    L.check(lib.cqb_memcpy_h2d(d, a.ctypes.data_as(ctypes.c_void_p), a.nbytes))


def _structure(a):
    """structured values among the random ones: zero, one, r - 1, a run of zeros"""
    a[0] = 0
    a[1] = P.int_to_limbs(P.MONT % P.R_MOD)
    a[2] = P.int_to_limbs(P.to_mont(P.R_MOD - 1, P.R_MOD))
    a[1000:1100] = 0
    a[-1] = P.int_to_limbs(P.to_mont(P.R_MOD - 1, P.R_MOD))


def _first_diff(got, exp):
    bad = np.nonzero((got != exp).any(axis=1))[0]
    return f"{bad.size} of {got.shape[0]} elements differ, first at {bad[:4]}" if bad.size else "equal"


# ------------------------------------------------------------------------------------------------------------------- NTT
@pytest.mark.parametrize("log_n", [20, 22, 24, 25, 26])
def test_best_fft_limb_exact_large(env, oracle, log_n):
    """forward transform vs the oracle's best_fft; 25 and 26 take the 4-pass schedule"""
    cq, L, lib = env
    n = 1 << log_n
    dv = Dev(L, lib)
    try:
        d_a = dv.alloc(n * 32)
        L.check(lib.cqb_synth_scalars_dev(0x5EED0002 + log_n, 0, n, d_a))
        a = _d2h(L, lib, d_a, (n, 4))
        _structure(a)
        _h2d(L, lib, d_a, a)
        dom = cq.EvaluationDomain(1, log_n)
        L.check(lib.cqb_ntt_bn254_fr_dev(d_a, L.p64(dom.omega), log_n))
        got = _d2h(L, lib, d_a, (n, 4))
        exp = oracle.best_fft(a, dom.omega, log_n, oracle.hw_threads())
        assert np.array_equal(got, exp), _first_diff(got, exp)
        if log_n in (22, 25):  # inverse (omega^-1, 1/n fused) back to the input, and against the oracle's ifft
            L.check(lib.cqb_intt_bn254_fr_dev(d_a, L.p64(dom.omega_inv), L.p64(dom.ifft_divisor), log_n))
            back = _d2h(L, lib, d_a, (n, 4))
            assert np.array_equal(back, a), _first_diff(back, a)
            exp_i = oracle.ifft(exp, dom.omega_inv, log_n, dom.ifft_divisor, oracle.hw_threads())
            assert np.array_equal(back, exp_i)
    finally:
        dv.free()


@pytest.mark.parametrize("log_n", [21, 25])
def test_host_pointer_ntt_limb_exact_large(env, oracle, log_n):
    """the drop-in entry point (host buffer, in place) at sizes that take the 3- and 4-pass schedules"""
    cq, L, lib = env
    n = 1 << log_n
    a = oracle.synth_scalars(0x5EED0102 + log_n, n)
    dom = cq.EvaluationDomain(1, log_n)
    exp = oracle.ifft(a, dom.omega_inv, log_n, dom.ifft_divisor, oracle.hw_threads())
    got = a.copy()
    cq.EvaluationDomain.ifft(got, dom.omega_inv, log_n, dom.ifft_divisor)
    assert np.array_equal(got, exp), _first_diff(got, exp)


@pytest.mark.parametrize("j,k", [(3, 19), (5, 20), (5, 22), (3, 24), (5, 24)])
def test_domain_wrappers_limb_exact_large(env, oracle, j, k):
    """lagrange_to_coeff, coeff_to_extended (n -> 2n and n -> 4n), divide_by_vanishing_poly + extended_to_coeff on the
    device-pointer entry points, limb for limb (poly/domain.rs:238-338); (3,24) and (5,24) reach the 4-pass schedule"""
    cq, L, lib = env
    od = oracle.domain_new(j, k)
    d = cq.EvaluationDomain(j, k)
    n, ne = 1 << k, 1 << d.extended_k
    th = oracle.hw_threads()
    dv = Dev(L, lib)
    try:
        d_a, d_e = dv.alloc(n * 32), dv.alloc(ne * 32)
        L.check(lib.cqb_synth_scalars_dev(0xAB00 + 8 * k + j, 0, n, d_a))
        a = _d2h(L, lib, d_a, (n, 4))
        _structure(a)
        _h2d(L, lib, d_a, a)
        # lagrange_to_coeff
        L.check(lib.cqb_intt_bn254_fr_dev(d_a, L.p64(d.omega_inv), L.p64(d.ifft_divisor), k))
        coeff = _d2h(L, lib, d_a, (n, 4))
        exp_coeff = oracle.lagrange_to_coeff(od, a, th)
        assert np.array_equal(coeff, exp_coeff), _first_diff(coeff, exp_coeff)
        # coeff_to_extended
        L.check(lib.cqb_coset_ntt_bn254_fr_dev(d_a, n, d_e, L.p64(d.extended_omega), d.extended_k, L.p64(d.g_coset), L.p64(d.g_coset_inv)))
        ext = _d2h(L, lib, d_e, (ne, 4))
        exp_ext = oracle.coeff_to_extended(od, exp_coeff, th)
        assert np.array_equal(ext, exp_ext), _first_diff(ext, exp_ext)
        del ext
        # quotient path on a random extended vector: divide_by_vanishing_poly -> extended_to_coeff (vanishing/prover.rs:84-87)
        L.check(lib.cqb_synth_scalars_dev(0xCD00 + 8 * k + j, 0, ne, d_e))
        h = _d2h(L, lib, d_e, (ne, 4))
        L.check(lib.cqb_coset_intt_bn254_fr_dev(d_e, d.extended_k, L.p64(d.extended_omega_inv), L.p64(d.extended_ifft_divisor), L.p64(d.g_coset),
                                                L.p64(d.g_coset_inv), L.p64(d.t_evaluations), d.t_evaluations.shape[0]))
        got = _d2h(L, lib, d_e, (ne, 4))[: n * d.quotient_poly_degree]
        exp = oracle.extended_to_coeff(od, oracle.divide_by_vanishing_poly(od, h), th)
        assert np.array_equal(got, exp), _first_diff(got, exp)
    finally:
        dv.free()


# ------------------------------------------------------------------------------------------------------------------- MSM
def _msm_dev(L, lib, h, d_s, n, offset=0):
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    L.check(lib.cqb_msm_bn254_g1_dev(h, offset, d_s, n, L.p64(out), ctypes.byref(inf)))
    return out


def _full_msm_vs_oracle(env, oracle, log_n, host_paths):
    cq, L, lib = env
    n = 1 << log_n
    dv = Dev(L, lib)
    hp = ctypes.c_void_p()
    try:
        d_b, d_s = dv.alloc(n * 64), dv.alloc(n * 32)
        L.check(lib.cqb_synth_bases_dev(0xC0FFEE, 0, n, d_b))
        L.check(lib.cqb_synth_scalars_dev(0x5EED0001, 0, n, d_s))
        sc = _d2h(L, lib, d_s, (n, 4))
        bs = _d2h(L, lib, d_b, (n, 8))
        _, exp = oracle.best_multiexp(sc, bs, oracle.hw_threads())
        del bs
        # the point bench.py asserts at every GPU count (tests/golden/bench_points.json) is this oracle result
        import json
        import os

        gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bench_points.json"))).get(str(log_n))
        if gold is not None:  # x, y as the 256-bit integers the in-memory G1Affine limbs spell (Montgomery form)
            ex = sum(int(v) << (64 * i) for i, v in enumerate(exp[:4]))
            ey = sum(int(v) << (64 * i) for i, v in enumerate(exp[4:]))
            assert (int(gold["x"], 16), int(gold["y"], 16)) == (ex, ey), "committed bench point differs from the oracle"
        h = ctypes.c_uint64(0)
        L.check(lib.cqb_bases_register_device(d_b, n, ctypes.byref(h)))
        try:
            got_w = _msm_dev(L, lib, h.value, d_s, n)  # windowed layout (c = 16)
            assert np.array_equal(got_w, exp), "windowed layout differs from the oracle"
            L.check(lib.cqb_bases_precompute(h.value, 0))  # single bucket set over the table (c = 20 at these sizes)
            assert lib.cqb_bases_precomputed_window_bits(h.value) >= 16
            got_t = _msm_dev(L, lib, h.value, d_s, n)
            assert np.array_equal(got_t, exp), "table layout differs from the oracle"
            # at these sizes the automatic choice is the affine tree (>= 40 entries per bucket): that is the path just checked ...
            assert lib.cqb_msm_last_tree_levels() >= 2, "the affine-tree accumulation was expected here"
            # ... and the XYZZ accumulation it replaces gives the same point
            L.check(lib.cqb_msm_set_accumulator(1, 0))
            try:
                got_x = _msm_dev(L, lib, h.value, d_s, n)
                assert lib.cqb_msm_last_tree_levels() == 0
            finally:
                L.check(lib.cqb_msm_set_accumulator(0, 0))
            assert np.array_equal(got_x, exp), "table layout with XYZZ accumulation differs from the oracle"
            if host_paths:
                out = np.zeros(8, np.uint64)
                inf = ctypes.c_int(0)
                # pageable host memory (what a Rust Vec<Fr> is)
                L.check(lib.cqb_msm_bn254_g1(h.value, 0, L.p64(sc), n, L.p64(out), ctypes.byref(inf)))
                assert np.array_equal(out, exp), "pageable host-pointer path differs from the oracle"
                # pinned host memory: the 3-part copy/compute pipeline
                L.check(lib.cqb_host_alloc_pinned(n * 32, ctypes.byref(hp)))
                ctypes.memmove(hp, sc.ctypes.data_as(ctypes.c_void_p), n * 32)
                out[:] = 0
                L.check(lib.cqb_msm_bn254_g1(h.value, 0, ctypes.cast(hp, L.u64p), n, L.p64(out), ctypes.byref(inf)))
                assert np.array_equal(out, exp), "pinned 3-part host-pointer path differs from the oracle"
        finally:
            L.check(lib.cqb_bases_free(h.value))
    finally:
        if hp.value:
            L.check(lib.cqb_host_free_pinned(hp))
        dv.free()


def test_msm_2p24_full_vs_oracle(env, oracle):
    """the headline size, every point: c = 20 / 2^19 buckets, the two-pass scatter and 128-entry chunks only occur here"""
    _full_msm_vs_oracle(env, oracle, 24, host_paths=True)


def test_msm_2p23_full_vs_oracle(env, oracle):
    _full_msm_vs_oracle(env, oracle, 23, host_paths=True)


@pytest.mark.slow
def test_msm_2p26_full_vs_oracle(env, oracle):
    """top of the sweep (about a minute of CPU for the oracle)"""
    _full_msm_vs_oracle(env, oracle, 26, host_paths=False)


@pytest.mark.slow
def test_commit_identity_at_k26_both_srs_vectors(env, oracle):
    """BASELINE.json configs[4]'s "largest k": ParamsKZG at k = 26 generated on the device (g and g_lagrange: 8 GiB), BOTH per-SRS
    tables resident (2 x 52 GiB at c = 20 — DESIGN.md section 2's memory plan), and the reference's own commitment identity
    (poly/kzg/commitment.rs:570-593 test_commit_lagrange): commit(lagrange_to_coeff(a)) == commit_lagrange(a), on 2^26 points."""
    cq, L, lib = env
    k = 26
    n = 1 << k
    s = oracle.synth_scalars(0x26, 1)[0]
    params = cq.ParamsKZG.setup_from_toxic_waste(k, s, precompute=True)
    dv = Dev(L, lib)
    try:
        assert lib.cqb_bases_precomputed_window_bits(params.g.handle) >= 16, "the table of g did not fit"
        assert lib.cqb_bases_precomputed_window_bits(params.g_lagrange.handle) >= 16, "the table of g_lagrange did not fit"
        d_a = dv.alloc(n * 32)
        L.check(lib.cqb_synth_scalars_dev(0x2626, 0, n, d_a))
        c_lagrange = _msm_dev(L, lib, params.g_lagrange.handle, d_a, n)
        dom = cq.EvaluationDomain(1, k)
        L.check(lib.cqb_intt_bn254_fr_dev(d_a, L.p64(dom.omega_inv), L.p64(dom.ifft_divisor), k))
        c_coeff = _msm_dev(L, lib, params.g.handle, d_a, n)
        assert c_lagrange.any() and np.array_equal(c_coeff, c_lagrange)
        # anchor to the oracle on a prefix: the first 2^16 coefficients against the first 2^16 powers of the SRS
        m = 1 << 16
        sc = _d2h(L, lib, d_a, (m, 4))
        g_host = np.zeros((m, 8), np.uint64)
        L.check(lib.cqb_bases_download(params.g.handle, 0, m, L.p64(g_host)))
        _, exp = oracle.best_multiexp(sc, g_host, oracle.hw_threads())
        # a short commitment (2^16 of the 2^26 points) over the same resident table
        assert np.array_equal(_msm_dev(L, lib, params.g.handle, d_a, m), exp)
    finally:
        dv.free()
        params.free()

"""evaluate_h's row-wise terms on the device vs the oracle (plonk/evaluation.rs:285-551):
GraphEvaluator interpreter (custom gates), the CQ static-lookup term and the permutation terms."""
import ctypes
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402
from tests.evalh_common import random_expr  # noqa: E402


@pytest.fixture(scope="module")
def cq():
    import cqb200

    cqb200._lib.init(0)
    return cqb200


class Dev:
    def __init__(self, cq):
        self.cq, self.lib, self.allocs = cq, cq._lib.lib(), []

    def up(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        d = ctypes.c_void_p()
        self.cq._lib.check(self.lib.cqb_dev_alloc(max(arr.nbytes, 64), ctypes.byref(d)))
        self.cq._lib.check(self.lib.cqb_memcpy_h2d(d, arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes))
        self.allocs.append(d)
        return d.value

    def down(self, ptr, n):
        out = np.zeros((n, 4), np.uint64)
        self.cq._lib.check(self.lib.cqb_memcpy_d2h(out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(ptr), out.nbytes))
        return out

    def free(self):
        for d in self.allocs:
            self.cq._lib.check(self.lib.cqb_dev_free(d))


@pytest.mark.parametrize("size,depth,npolys", [(16, 3, 1), (256, 5, 4), (4096, 6, 8), (1 << 15, 7, 24)])
def test_graph_evaluate_parity(cq, oracle, size, depth, npolys):
    from sha2_on_cq_halo2_b200.evaluation import Expr, custom_gates_evaluator

    rng = random.Random(size)
    polys = [random_expr(rng, Expr, depth, ncols=(3, 5, 2), nchal=3) for _ in range(npolys)]
    ev = custom_gates_evaluator(polys)
    fixed = [oracle.synth_scalars(100 + i, size) for i in range(3)]
    advice = [oracle.synth_scalars(200 + i, size) for i in range(5)]
    inst = [oracle.synth_scalars(300 + i, size) for i in range(2)]
    chal = oracle.synth_scalars(400, 3)
    beta, gamma, theta, y = oracle.synth_scalars(500, 4)
    prev = oracle.synth_scalars(600, size)
    consts, rots, code = ev.serialize()
    exp = oracle.graph_evaluate(consts, rots, code, len(ev.calculations), ev.num_intermediates, fixed, advice, inst, chal, beta, gamma,
                                theta, y, prev, 4)
    dv = Dev(cq)
    d_vals = dv.up(prev)
    ev.evaluate_dev([dv.up(c) for c in fixed], [dv.up(c) for c in advice], [dv.up(c) for c in inst], chal, beta, gamma, theta, y, d_vals,
                    size, 4)
    assert np.array_equal(dv.down(d_vals, size), exp), (ev.num_intermediates, len(ev.calculations))
    dv.free()


def test_cq_lookup_term_parity(cq, oracle):
    from sha2_on_cq_halo2_b200.evaluation import cq_lookup_h_dev

    size = 5000
    v, b, f, l = (oracle.synth_scalars(700 + i, size) for i in range(4))
    beta, y = oracle.synth_scalars(710, 2)
    dv = Dev(cq)
    d_v = dv.up(v)
    cq_lookup_h_dev(d_v, dv.up(b), dv.up(f), dv.up(l), beta, y, size)
    assert np.array_equal(dv.down(d_v, size), oracle.cq_lookup_h(v, b, f, l, beta, y))
    dv.free()


@pytest.mark.parametrize("nsets,ncols,chunk_len", [(1, 1, 1), (2, 5, 3), (3, 7, 3)])
def test_permutation_terms_parity(cq, oracle, nsets, ncols, chunk_len):
    from sha2_on_cq_halo2_b200.evaluation import permutation_h_dev

    size, rot_scale, last_rotation = 2048, 2, -6
    v = oracle.synth_scalars(800, size)
    sets = [oracle.synth_scalars(810 + i, size) for i in range(nsets)]
    cols = [oracle.synth_scalars(820 + i, size) for i in range(ncols)]
    perms = [oracle.synth_scalars(840 + i, size) for i in range(ncols)]
    l0, l_last, l_act = (oracle.synth_scalars(860 + i, size) for i in range(3))
    beta, gamma, y = oracle.synth_scalars(870, 3)
    ew = P.int_to_limbs(P.to_mont(P.omega_for(11), P.R_MOD))
    exp = oracle.permutation_h(v, rot_scale, last_rotation, chunk_len, sets, cols, perms, l0, l_last, l_act, beta, gamma, y, ew)
    dv = Dev(cq)
    d_v = dv.up(v)
    permutation_h_dev(d_v, size, rot_scale, last_rotation, chunk_len, [dv.up(s) for s in sets], [dv.up(c) for c in cols],
                      [dv.up(p) for p in perms], dv.up(l0), dv.up(l_last), dv.up(l_act), beta, gamma, y, ew)
    assert np.array_equal(dv.down(d_v, size), exp)
    dv.free()


def test_quotient_pipeline_device_resident(cq, oracle):
    """The chain evaluate_h sits in (plonk/prover.rs:606-627): coeff_to_extended of the advice / CQ polynomials (a8) ->
    custom gates + CQ term row by row (evaluate_h) -> divide_by_vanishing_poly + extended_to_coeff (a9/a10) -> commit the
    h pieces (a16) — every intermediate stays in HBM; only the commitments come back. Compared with the oracle doing the
    same steps on the host."""
    from sha2_on_cq_halo2_b200.evaluation import Expr, cq_lookup_h_dev, custom_gates_evaluator

    L, lib = cq._lib, cq._lib.lib()
    k = 6
    n = 1 << k
    dom = cq.EvaluationDomain(3, k)
    od = oracle.domain_new(3, k)
    en = 1 << dom.extended_k
    rot_scale = 1 << (dom.extended_k - k)
    s = oracle.synth_scalars(0xAB, 1)[0]
    g, gl = oracle.params_setup(k, s)
    params = cq.ParamsKZG(k, g, gl, precompute=False)
    # coefficient-form polynomials: two advice columns, one fixed (selector-like), CQ's b and f, and l_active_row
    polys = {name: oracle.synth_scalars(0x900 + i, n) for i, name in enumerate(["a0", "a1", "q", "b", "f", "lact"])}
    beta, gamma, theta, y = oracle.synth_scalars(0x950, 4)
    # gate: q * (a0 * a1(rot 1) - a1) ; q * (a0 - a0(rot -1)) * 7
    a0, a1, a1n, a0p, q = Expr("advice", 0, 0), Expr("advice", 1, 0), Expr("advice", 1, 1), Expr("advice", 0, -1), Expr("fixed", 0, 0)
    ev = custom_gates_evaluator([q * (a0 * a1n - a1), (q * (a0 - a0p)) * 7])
    consts, rots, code = ev.serialize()
    # ---- oracle ----
    ext = {name: oracle.coeff_to_extended(od, p) for name, p in polys.items()}
    h = oracle.graph_evaluate(consts, rots, code, len(ev.calculations), ev.num_intermediates, [ext["q"]], [ext["a0"], ext["a1"]], [],
                              np.zeros((0, 4), np.uint64), beta, gamma, theta, y, np.zeros((en, 4), np.uint64), rot_scale)
    h = oracle.cq_lookup_h(h, ext["b"], ext["f"], ext["lact"], beta, y)
    h_coeff = oracle.extended_to_coeff(od, oracle.divide_by_vanishing_poly(od, h))
    exp_cms = [oracle.best_multiexp(np.ascontiguousarray(h_coeff[i * n:(i + 1) * n]), g, 2)[1] for i in range(dom.quotient_poly_degree)]
    # ---- device-resident ----
    def dalloc(nbytes):
        d = ctypes.c_void_p()
        L.check(lib.cqb_dev_alloc(nbytes, ctypes.byref(d)))
        return d

    d_ext = {}
    d_coeff = dalloc(n * 32)
    for name, p in polys.items():
        L.check(lib.cqb_memcpy_h2d(d_coeff, p.ctypes.data_as(ctypes.c_void_p), n * 32))
        d_ext[name] = dalloc(en * 32)
        L.check(lib.cqb_coset_ntt_bn254_fr_dev(d_coeff, n, d_ext[name], L.p64(dom.extended_omega), dom.extended_k, L.p64(dom.g_coset),
                                               L.p64(dom.g_coset_inv)))
    d_h = dalloc(en * 32)
    L.check(lib.cqb_memcpy_h2d(d_h, np.zeros((en, 4), np.uint64).ctypes.data_as(ctypes.c_void_p), en * 32))
    ev.evaluate_dev([d_ext["q"].value], [d_ext["a0"].value, d_ext["a1"].value], [], [], beta, gamma, theta, y, d_h.value, en, rot_scale)
    cq_lookup_h_dev(d_h.value, d_ext["b"].value, d_ext["f"].value, d_ext["lact"].value, beta, y, en)
    L.check(lib.cqb_coset_intt_bn254_fr_dev(d_h, dom.extended_k, L.p64(dom.extended_omega_inv), L.p64(dom.extended_ifft_divisor),
                                            L.p64(dom.g_coset), L.p64(dom.g_coset_inv), L.p64(dom.t_evaluations), dom.t_evaluations.shape[0]))
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    for i in range(dom.quotient_poly_degree):
        L.check(lib.cqb_msm_bn254_g1_dev(params.g.handle, 0, ctypes.c_void_p(d_h.value + i * n * 32), n, L.p64(out), ctypes.byref(inf)))
        assert np.array_equal(out, exp_cms[i]), i
    got_h = np.zeros((n * dom.quotient_poly_degree, 4), np.uint64)
    L.check(lib.cqb_memcpy_d2h(got_h.ctypes.data_as(ctypes.c_void_p), d_h, got_h.nbytes))
    assert np.array_equal(got_h, h_coeff)
    for d in list(d_ext.values()) + [d_coeff, d_h]:
        L.check(lib.cqb_dev_free(d))
    params.free()

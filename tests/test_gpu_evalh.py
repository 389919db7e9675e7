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

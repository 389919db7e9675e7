"""CPU: the GraphEvaluator builder mirror (plonk/evaluation.rs:571-704) + the oracle's interpreter (:718-775) reproduce
the direct evaluation of the expressions (:778-816) combined by the custom-gate Horner (:228-246)."""
import random

import numpy as np

from oracle import pyref as P
from tests.evalh_common import eval_expr, random_expr


def test_graph_evaluator_matches_direct_expression_evaluation(oracle):
    import cqb200
    from sha2_on_cq_halo2_b200.evaluation import Expr, custom_gates_evaluator

    rng = random.Random(42)
    size, rot_scale = 32, 2
    for trial in range(12):
        cols = {k: [[rng.randrange(P.R_MOD) for _ in range(size)] for _ in range(n)] for k, n in (("f", 2), ("a", 3), ("i", 1))}
        chal = [rng.randrange(P.R_MOD) for _ in range(2)]
        polys = [random_expr(rng, Expr, 4) for _ in range(rng.randint(1, 4))]
        ev = custom_gates_evaluator(polys)
        consts, rots, code = ev.serialize()
        beta, gamma, theta, y = (rng.randrange(P.R_MOD) for _ in range(4))
        prev = [rng.randrange(P.R_MOD) for _ in range(size)]
        L = lambda v: P.int_to_limbs(P.to_mont(v, P.R_MOD))
        out = oracle.graph_evaluate(consts, rots, code, len(ev.calculations), ev.num_intermediates,
                                    [P.fr_array_from_ints(c) for c in cols["f"]], [P.fr_array_from_ints(c) for c in cols["a"]],
                                    [P.fr_array_from_ints(c) for c in cols["i"]], P.fr_array_from_ints(chal), L(beta), L(gamma), L(theta),
                                    L(y), P.fr_array_from_ints(prev), rot_scale)
        got = P.fr_array_to_ints(out)
        for idx in range(size):
            acc = prev[idx]  # Horner(PreviousValue, parts, y): value = value * y + part
            for p in polys:
                acc = (acc * y + eval_expr(p, idx, size, rot_scale, cols["f"], cols["a"], cols["i"], chal)) % P.R_MOD
            assert got[idx] == acc, (trial, idx)

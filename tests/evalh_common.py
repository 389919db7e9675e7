"""shared helpers for the evaluate_h tests: random expression trees and their direct (big-integer) evaluation — the
semantics of the reference's `evaluate` (plonk/evaluation.rs:778-816)."""
import random

from oracle import pyref as P


def random_expr(rng, E, depth, ncols=(2, 3, 1), nchal=2):
    if depth == 0 or rng.random() < 0.15:
        t = rng.choice(["const", "fixed", "advice", "advice", "instance", "challenge"])
        if t == "const":
            return E("const", rng.choice([0, 1, 2, 5, P.R_MOD - 1, rng.randrange(P.R_MOD)]))
        if t == "challenge":
            return E("challenge", rng.randrange(nchal))
        n = {"fixed": ncols[0], "advice": ncols[1], "instance": ncols[2]}[t]
        return E(t, rng.randrange(n), rng.choice([0, 0, 1, -1, 2, -3]))
    op = rng.choice(["sum", "sub", "prod", "prod", "neg", "scaled"])
    a = random_expr(rng, E, depth - 1, ncols, nchal)
    if op == "neg":
        return -a
    if op == "scaled":
        return a * rng.choice([0, 1, 2, 7, rng.randrange(P.R_MOD)])
    b = random_expr(rng, E, depth - 1, ncols, nchal) if rng.random() < 0.8 else a
    return {"sum": a + b, "sub": a - b, "prod": a * b}[op]


def eval_expr(e, idx, size, rot_scale, fixed, advice, instance, challenges):
    """fixed/advice/instance: lists of lists of canonical ints"""
    n = e.node
    t = n[0]
    if t == "const":
        return n[1] % P.R_MOD
    if t in ("fixed", "advice", "instance"):
        col = {"fixed": fixed, "advice": advice, "instance": instance}[t][n[1]]
        return col[(idx + n[2] * rot_scale) % size]
    if t == "challenge":
        return challenges[n[1]]
    if t == "neg":
        return (-eval_expr(n[1], idx, size, rot_scale, fixed, advice, instance, challenges)) % P.R_MOD
    if t == "scaled":
        return eval_expr(n[1], idx, size, rot_scale, fixed, advice, instance, challenges) * n[2] % P.R_MOD
    a = eval_expr(n[1], idx, size, rot_scale, fixed, advice, instance, challenges)
    b = eval_expr(n[2], idx, size, rot_scale, fixed, advice, instance, challenges)
    return (a + b) % P.R_MOD if t == "sum" else a * b % P.R_MOD

"""tools/bench_next_rows.py — timings of the SURVEY.md §8(f) rows built beyond the MSM/NTT core (one B200, wall clock
around synchronous calls after a warm-up): SRS generation, g_to_lagrange, FK table preprocessing, element-wise helpers.
The CPU column is the oracle's restatement of the reference at a size it finishes in seconds (scaled where noted)."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import cqb200
from oracle import oracle_lib as O

L = cqb200._lib
lib = L.init(0)
res = {}
s = O.synth_scalars(1, 1)[0]
threads = O.hw_threads()


def wall(fn, reps=2):
    fn()
    L.check(lib.cqb_sync())
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    L.check(lib.cqb_sync())
    return (time.perf_counter() - t) / reps * 1e3


def dev(nbytes):
    d = ctypes.c_void_p()
    L.check(lib.cqb_dev_alloc(nbytes, ctypes.byref(d)))
    return d


# SRS generation (commitment.rs:209-276)
for k in (16, 20, 24):
    d = dev(2 * (1 << k) * 64)
    ms = wall(lambda: L.check(lib.cqb_srs_setup_dev(k, L.p64(s), d, ctypes.c_void_p(d.value + (1 << k) * 64))))
    res[f"srs_setup_k{k}_ms"] = round(ms, 2)
    L.check(lib.cqb_dev_free(d))
t = time.perf_counter()
O.params_setup(8, s)
res["cpu_srs_setup_k8_ms_1thread"] = round((time.perf_counter() - t) * 1e3, 1)

# g_to_lagrange (arithmetic.rs:277-301)
for k in (12, 16, 18):
    n = 1 << k
    d = dev(2 * n * 64)
    L.check(lib.cqb_synth_bases_dev(7, 0, n, d))
    ms = wall(lambda: L.check(lib.cqb_g_to_lagrange_dev(d, k, ctypes.c_void_p(d.value + n * 64))), reps=1)
    res[f"g_to_lagrange_k{k}_ms"] = round(ms, 2)
    L.check(lib.cqb_dev_free(d))
g8 = O.synth_bases(7, 256, threads)
t = time.perf_counter()
O.g_to_lagrange(g8, 8)
res["cpu_g_to_lagrange_k8_ms_1thread"] = round((time.perf_counter() - t) * 1e3, 1)

# FK table preprocessing (static_lookup.rs:77-126)
for k in (12, 16):
    n = 1 << k
    d = dev(n * 32 + 2 * n * 64)
    L.check(lib.cqb_synth_scalars_dev(9, 0, n, d))
    d_srs = ctypes.c_void_p(d.value + n * 32)
    d_qs = ctypes.c_void_p(d.value + n * 32 + n * 64)
    L.check(lib.cqb_synth_bases_dev(11, 0, n, d_srs))
    ms = wall(lambda: L.check(lib.cqb_cq_table_qs_dev(d, k, d_srs, d_qs)), reps=1)
    res[f"cq_table_qs_fk_N2^{k}_ms"] = round(ms, 2)
    L.check(lib.cqb_dev_free(d))
vals = O.synth_scalars(9, 256)
srs = O.synth_bases(11, 256, threads)
t = time.perf_counter()
O.cq_table_qs(vals, srs, threads)
cpu_ms = (time.perf_counter() - t) * 1e3
res["cpu_cq_table_qs_N2^8_ms"] = round(cpu_ms, 1)
res["cpu_cq_table_qs_N2^16_estimate_s"] = round(cpu_ms * (65536 / 256) ** 2 / 1e3, 0)
res["cpu_threads"] = threads

# element-wise helpers at 2^24
n = 1 << 24
d = dev(n * 32)
dq = dev(n * 32)
L.check(lib.cqb_synth_scalars_dev(13, 0, n, d))
x = O.synth_scalars(14, 1)[0]
out = np.zeros(4, np.uint64)
res["eval_polynomial_2^24_ms"] = round(wall(lambda: L.check(lib.cqb_eval_polynomial_dev(d, n, L.p64(x), L.p64(out)))), 3)
res["kate_division_2^24_ms"] = round(wall(lambda: L.check(lib.cqb_kate_division_dev(d, n, L.p64(x), dq))), 3)
res["batch_invert_2^24_ms"] = round(wall(lambda: L.check(lib.cqb_fr_batch_invert_dev(d, n))), 3)
res["prefix_product_2^24_ms"] = round(wall(lambda: L.check(lib.cqb_fr_prefix_product_dev(d, n, L.p64(x), dq))), 3)
# permutation grand product (permutation/prover.rs:82-166): one set of 4 columns at k = 20 and k = 24
from sha2_on_cq_halo2_b200.permutation import product_set_dev
from oracle import pyref as P
for kk in (20, 24):
    nn = 1 << kk
    pcols = [dev(nn * 32) for _ in range(8)]
    for i, c in enumerate(pcols):
        L.check(lib.cqb_synth_scalars_dev(700 + i, 0, nn, c))
    res[f"permutation_product_4cols_k{kk}_ms"] = round(wall(lambda: product_set_dev([c.value for c in pcols[:4]], [c.value for c in pcols[4:]], kk, 5, 7,
                                                                                P.omega_for(kk), 1, 1, dq.value if kk == 24 else pcols[0].value)), 3)
    for c in pcols:
        L.check(lib.cqb_dev_free(c))
pc = [O.synth_scalars(700 + i, 1 << 16) for i in range(8)]
one = P.fr_array_from_ints([1])[0]
t = time.perf_counter()
O.permutation_product(pc[:4], pc[4:], P.fr_array_from_ints([5])[0], P.fr_array_from_ints([7])[0], P.fr_array_from_ints([P.omega_for(16)])[0], one, one)
res["cpu_permutation_product_4cols_k16_ms_1thread"] = round((time.perf_counter() - t) * 1e3, 1)
a20 = O.synth_scalars(13, 1 << 20)
t = time.perf_counter()
O.kate_division(a20, x)
res["cpu_kate_division_2^20_ms_1thread"] = round((time.perf_counter() - t) * 1e3, 1)
print(json.dumps(res))

# evaluate_h (plonk/evaluation.rs:285-551): custom-gate graph over 2^22 extended rows + CQ term
import random

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.evalh_common import random_expr
from sha2_on_cq_halo2_b200.evaluation import Expr, cq_lookup_h_dev, custom_gates_evaluator

rng = random.Random(5)
polys = [random_expr(rng, Expr, 6, ncols=(4, 8, 1), nchal=2) for _ in range(16)]
ev = custom_gates_evaluator(polys)
size = 1 << 22
cols = []
for i in range(13):
    dcol = dev(size * 32)
    L.check(lib.cqb_synth_scalars_dev(100 + i, 0, size, dcol))
    cols.append(dcol)
d_vals = dev(size * 32)
L.check(lib.cqb_synth_scalars_dev(99, 0, size, d_vals))
chal = O.synth_scalars(400, 2)
beta, gamma, theta, y = O.synth_scalars(500, 4)
fixed, advice, inst = [c.value for c in cols[:4]], [c.value for c in cols[4:12]], [cols[12].value]
ms = wall(lambda: ev.evaluate_dev(fixed, advice, inst, chal, beta, gamma, theta, y, d_vals.value, size, 2))
nmul = sum(1 for c, _ in ev.calculations if c[0] in (2, 3)) + sum(len(c[2]) for c, _ in ev.calculations if c[0] == 6)
res["evaluate_h_graph_2^22_rows_ms"] = round(ms, 3)
res["evaluate_h_graph_calculations"] = len(ev.calculations)
res["evaluate_h_graph_intermediates"] = ev.num_intermediates
res["evaluate_h_graph_modmuls_per_row"] = nmul
res["evaluate_h_graph_Gmodmul_per_s"] = round(nmul * size / ms / 1e6, 2)
res["cq_lookup_term_2^22_rows_ms"] = round(wall(lambda: cq_lookup_h_dev(d_vals.value, cols[0].value, cols[1].value, cols[2].value, beta, y, size)), 3)
small = 1 << 14
c_small = [O.synth_scalars(100 + i, small) for i in range(13)]
consts, rots, code = ev.serialize()
t = time.perf_counter()
O.graph_evaluate(consts, rots, code, len(ev.calculations), ev.num_intermediates, c_small[:4], c_small[4:12], c_small[12:], chal, beta, gamma,
                 theta, y, O.synth_scalars(99, small), 2)
res["cpu_evaluate_h_graph_2^14_rows_ms_1thread"] = round((time.perf_counter() - t) * 1e3, 1)
print(json.dumps(res))

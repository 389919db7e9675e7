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
a20 = O.synth_scalars(13, 1 << 20)
t = time.perf_counter()
O.kate_division(a20, x)
res["cpu_kate_division_2^20_ms_1thread"] = round((time.perf_counter() - t) * 1e3, 1)
print(json.dumps(res))

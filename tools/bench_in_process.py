"""tools/bench_in_process.py — the one-process-N-devices MSM (cqb_init_multi) timed stand-alone, no torch: whole set, and the halves /
shards alone, to see whether the devices overlap.   python tools/bench_in_process.py <n_devices> <log_n>"""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cqb200

L = cqb200._lib
lib = L.load()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
n = 1 << log_n
L.check(lib.cqb_init_multi(G))
d = ctypes.c_void_p()
bases_h = np.empty((n, 8), np.uint64)
L.check(lib.cqb_dev_alloc(n * 64, ctypes.byref(d)))
L.check(lib.cqb_synth_bases_dev(0xC0FFEE, 0, n, d))
L.check(lib.cqb_memcpy_d2h(bases_h.ctypes.data_as(ctypes.c_void_p), d, n * 64))
L.check(lib.cqb_dev_free(d))
h = ctypes.c_uint64(0)
L.check(lib.cqb_bases_register_sharded(L.p64(bases_h), n, ctypes.byref(h)))
L.check(lib.cqb_bases_precompute(h.value, 0))
print("table c =", lib.cqb_bases_precomputed_window_bits(h.value), flush=True)
ptrs = (ctypes.c_void_p * G)()
base, rem = divmod(n, G)
starts = []
for i in range(G):
    s0, c0 = i * base + min(i, rem), base + (1 if i < rem else 0)
    starts.append((s0, c0))
    p = ctypes.c_void_p()
    L.check(lib.cqb_dev_alloc_on(i, c0 * 32, ctypes.byref(p)))
    L.check(lib.cqb_synth_scalars_dev_on(i, 0x5EED0001, s0, c0, p))
    ptrs[i] = p.value
out, inf = np.zeros(8, np.uint64), ctypes.c_int(0)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


res = {"n_devices": G, "log_n": log_n}
res["all_ms"] = timed(lambda: L.check(lib.cqb_msm_bn254_g1_multi_dev(h.value, 0, ptrs, n, L.p64(out), ctypes.byref(inf))))
res["point_x0"] = hex(int(out[0]))
for i, (s0, c0) in enumerate(starts[:4]):
    one = (ctypes.c_void_p * G)()
    one[i] = ptrs[i]
    res[f"shard{i}_only_ms"] = timed(lambda: L.check(lib.cqb_msm_bn254_g1_multi_dev(h.value, s0, one, c0, L.p64(out), ctypes.byref(inf))))
print(json.dumps(res), flush=True)

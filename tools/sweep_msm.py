"""tools/sweep_msm.py — GPU-side tuning sweep: MSM time vs n and window bits (device-resident inputs). Prints JSON lines."""
import ctypes
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cqb200

L = cqb200._lib
lib = L.init(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
L.check(lib.cqb_set_stream(ctypes.c_void_p(stream.cuda_stream)))
L.check(lib.cqb_msm_set_profiling(1))
L.check(lib.cqb_msm_set_parts(int(os.environ.get("CQB_PARTS", "0"))))  # 0 = automatic
_acc = [int(x) for x in os.environ.get("CQB_ACC", "0,0").split(",")]  # accumulation variant, affine segment log
L.check(lib.cqb_msm_set_accumulator(_acc[0], _acc[1]))
logs = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "16,18,20,22,24".split(","))]
cs = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0".split(","))]
nmax = 1 << max(logs)
bases = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
scal = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
L.check(lib.cqb_synth_bases_dev(0xC0FFEE, 0, nmax, ctypes.c_void_p(bases.data_ptr())))
L.check(lib.cqb_synth_scalars_dev(0x5EED0001, 0, nmax, ctypes.c_void_p(scal.data_ptr())))
dist = os.environ.get("CQB_DIST", "")  # optional skewed scalar distribution (same definitions as tools/sweep_all.py)
if dist:
    from oracle import pyref as P
    rng = np.random.default_rng(7)
    full = np.zeros((nmax, 4), np.uint64)
    L.check(lib.cqb_memcpy_d2h(full.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(scal.data_ptr()), nmax * 32))
    if dist == "negative_small":
        tab = P.fr_array_from_ints([P.R_MOD - v for v in range(1, 1 << 12)])
        arr = tab[rng.integers(0, (1 << 12) - 1, nmax)]
    elif dist == "all_equal":
        arr = np.repeat(full[:1], nmax, axis=0)
    elif dist == "two_values":
        arr = full[rng.integers(0, 2, nmax)]
    elif dist == "bits":
        arr = P.fr_array_from_ints([0, 1])[rng.integers(0, 2, nmax)]
    else:
        raise SystemExit("unknown CQB_DIST")
    arr = np.ascontiguousarray(arr)
    L.check(lib.cqb_memcpy_h2d(ctypes.c_void_p(scal.data_ptr()), arr.ctypes.data_as(ctypes.c_void_p), nmax * 32))
h = ctypes.c_uint64(0)
L.check(lib.cqb_bases_register_device(ctypes.c_void_p(bases.data_ptr()), nmax, ctypes.byref(h)))
out = np.zeros(8, np.uint64)
inf = ctypes.c_int(0)
ph = (ctypes.c_float * 8)()
pre = len(sys.argv) > 3 and sys.argv[3] == "pre"
for lg in logs:
    n = 1 << lg
    if pre:
        hh = ctypes.c_uint64(0)
        L.check(lib.cqb_bases_register_device(ctypes.c_void_p(bases.data_ptr()), n, ctypes.byref(hh)))
    for c in cs:
        if pre:
            import time
            t0 = time.time()
            L.check(lib.cqb_bases_precompute(hh.value, c))
            pre_s = time.time() - t0
            h = hh
        else:
            L.check(lib.cqb_msm_set_window_bits(c))
        def run():
            L.check(lib.cqb_msm_bn254_g1_dev(h.value, 0, ctypes.c_void_p(scal.data_ptr()), n, L.p64(out), ctypes.byref(inf)))
        for _ in range(2):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if lg <= 22 else 3
        e0.record(stream)
        for _ in range(reps):
            run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        lib.cqb_msm_phase_ms(ph, 8)
        print(json.dumps({"log_n": lg, "c": c, "ms": round(ms, 4), "mpts": round(n / ms / 1e3, 2),
                          "phases": [round(ph[i], 4) for i in range(8)], "x0": hex(int(out[0])), "acc": _acc, "precompute_s": round(pre_s, 3) if pre else None}), flush=True)
    if pre:
        L.check(lib.cqb_bases_free(hh.value))

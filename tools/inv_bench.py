import ctypes, json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import cqb200
L = cqb200._lib; lib = L.init(0)
def wall(fn, reps=3):
    fn(); L.check(lib.cqb_sync()); t = time.perf_counter()
    for _ in range(reps): fn()
    L.check(lib.cqb_sync()); return (time.perf_counter() - t) / reps * 1e3
for lg in (16, 20, 22, 24):
    n = 1 << lg
    d = ctypes.c_void_p(); L.check(lib.cqb_dev_alloc(n * 32, ctypes.byref(d)))
    L.check(lib.cqb_synth_scalars_dev(13, 0, n, d))
    print(lg, round(wall(lambda: L.check(lib.cqb_fr_batch_invert_dev(d, n))), 3), flush=True)
    L.check(lib.cqb_dev_free(d))

"""quick bring-up of the affine-tree accumulation: a few parity cases against the oracle, then timing against XYZZ"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cqb200
from oracle import oracle_lib as O
cqb200._lib.init(0)
L, lib = cqb200._lib, cqb200._lib.lib()
ok = True
for levels in (1, 2, 4, 5):
    L.check(lib.cqb_msm_set_tree_levels(levels))
    for n, c in ((700, 8), (5000, 8), (1 << 16, 11)):
        sc, bs = O.synth_scalars(3 + n, n), O.synth_bases(4 + n, n, 4)
        bs[5] = 0; bs[7] = bs[6]; sc[7] = sc[6]; bs[9] = O.g1_neg_a(bs[8]); sc[9] = sc[8]
        dev = cqb200.DeviceBases(bs, precompute=True, window_bits=c)
        exp = O.best_multiexp(sc, bs, 8)[1]
        L.check(lib.cqb_msm_set_accumulator(1, 0))
        a = dev.msm(sc).to_affine()
        L.check(lib.cqb_msm_set_accumulator(3, 0))
        b = dev.msm(sc).to_affine()
        sk = sc.copy(); sk[:] = sk[0]
        e2 = O.best_multiexp(sk, bs, 8)[1]
        b2 = dev.msm(sk).to_affine()
        good = np.array_equal(a, exp) and np.array_equal(b, exp) and np.array_equal(b2, e2)
        ok = ok and good
        print("levels", levels, "n", n, "c", c, "xyzz", np.array_equal(a, exp), "tree", np.array_equal(b, exp), "tree all-equal", np.array_equal(b2, e2), flush=True)
        dev.free()
print("PARITY", "OK" if ok else "FAIL", flush=True)

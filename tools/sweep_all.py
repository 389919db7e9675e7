"""tools/sweep_all.py — the size sweeps of BASELINE.json configs[1] and configs[2] on one B200 (device-resident inputs,
CUDA-event timing): MSM 2^16..2^26 (uniform scalars; skewed distributions reported separately) and NTT 2^16..2^26
(forward, inverse, coset n->2n and n->4n, coset inverse with vanishing division). Prints one JSON object."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cqb200
from sha2_on_cq_halo2_b200.fields import R_MOD

L = cqb200._lib
lib = L.init(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
L.check(lib.cqb_set_stream(ctypes.c_void_p(stream.cuda_stream)))
HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
INT_PEAK = 9.306  # profiles/INT_PEAK.json
max_log = int(sys.argv[1]) if len(sys.argv) > 1 else 26


def timed(fn, reps):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = np.zeros(8, np.uint64)
inf = ctypes.c_int(0)
res = {"msm_uniform": [], "msm_windowed": [], "msm_distributions": [], "ntt": []}

# ---------------------------------------------------------------------------------------------------------------- MSM
nmax = 1 << max_log
bases = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
scal = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
L.check(lib.cqb_synth_bases_dev(0xC0FFEE, 0, nmax, ctypes.c_void_p(bases.data_ptr())))
L.check(lib.cqb_synth_scalars_dev(0x5EED0001, 0, nmax, ctypes.c_void_p(scal.data_ptr())))
for lg in range(16, max_log + 1, 2):
    n = 1 << lg
    h = ctypes.c_uint64(0)
    L.check(lib.cqb_bases_register_device(ctypes.c_void_p(bases.data_ptr()), n, ctypes.byref(h)))

    def run():
        L.check(lib.cqb_msm_bn254_g1_dev(h.value, 0, ctypes.c_void_p(scal.data_ptr()), n, L.p64(out), ctypes.byref(inf)))

    reps = 10 if lg <= 20 else (5 if lg <= 24 else 3)
    ms_w = timed(run, reps)
    res["msm_windowed"].append({"log_n": lg, "ms": round(ms_w, 4), "mpts": round(n / ms_w / 1e3, 2)})
    L.check(lib.cqb_bases_precompute(h.value, 0))
    ms = timed(run, reps)
    # whole-MSM reading: the MAD32 the accumulation EXECUTES per point (bench.py mad32_per_entry) and, beside it, SURVEY's pinned 21,760
    lv = int(lib.cqb_msm_last_tree_levels())
    nwin = 254 // int(lib.cqb_bases_precomputed_window_bits(h.value)) + 1
    per_entry = 1232.0 if lv == 0 else (1 - 2.0 ** -lv) * 796.5 + 2.0 ** -lv * 1232
    res["msm_uniform"].append({"log_n": lg, "ms": round(ms, 4), "mpts": round(n / ms / 1e3, 2), "affine_tree_levels": lv,
                               "int_executed_frac": round(n * nwin * per_entry / (ms * 1e-3) / 1e12 / INT_PEAK, 3),
                               "int_pinned_algorithm_frac": round(n * 21760 / (ms * 1e-3) / 1e12 / INT_PEAK, 3)})
    if lg == 22:  # skewed scalar distributions, same bases
        rng = np.random.default_rng(7)
        full = np.zeros((n, 4), np.uint64)
        L.check(lib.cqb_memcpy_d2h(full.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(scal.data_ptr()), n * 32))
        R = (1 << 256) % R_MOD

        def mont_small(v):  # small ints -> Montgomery limbs, vectorised through Python ints on the few distinct values
            uniq, invm = np.unique(v, return_inverse=True)
            tab = np.array([[((int(u) * R % R_MOD) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for u in uniq], dtype=np.uint64)
            return tab[invm]

        dists = {
            "witness_like_90pct_zero_rest_lt_2^16": mont_small(np.where(rng.random(n) < 0.9, 0, rng.integers(0, 1 << 16, n))),
            "bits_0_1": mont_small(rng.integers(0, 2, n)),
            "all_equal": np.repeat(full[:1], n, axis=0),
            "negative_small_r_minus_lt_2^16": mont_small(np.array([R_MOD - int(v) for v in range(1, 1 << 12)], dtype=object)[rng.integers(0, (1 << 12) - 1, n)]),
            "two_values": full[rng.integers(0, 2, n)],
            "32_values": full[rng.integers(0, 32, n)],
            "small_range_lt_2^8": mont_small(rng.integers(0, 256, n)),
        }
        tmp = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        for name, arr in dists.items():
            arr = np.ascontiguousarray(arr)
            L.check(lib.cqb_memcpy_h2d(ctypes.c_void_p(tmp.data_ptr()), arr.ctypes.data_as(ctypes.c_void_p), n * 32))

            def run_d():
                L.check(lib.cqb_msm_bn254_g1_dev(h.value, 0, ctypes.c_void_p(tmp.data_ptr()), n, L.p64(out), ctypes.byref(inf)))

            msd = timed(run_d, 3)
            res["msm_distributions"].append({"log_n": lg, "dist": name, "ms": round(msd, 4), "mpts": round(n / msd / 1e3, 2)})
        del tmp
    L.check(lib.cqb_bases_free(h.value))
del bases
torch.cuda.empty_cache()

# ---------------------------------------------------------------------------------------------------------------- NTT
for k in range(16, max_log + 1, 2):
    n = 1 << k
    d1 = cqb200.EvaluationDomain(3, k)   # extended_k = k + 1
    d2 = cqb200.EvaluationDomain(5, k)   # extended_k = k + 2
    a = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    L.check(lib.cqb_synth_scalars_dev(0x5EED0002, 0, n, ctypes.c_void_p(a.data_ptr())))
    reps = 10 if k <= 22 else 4
    row = {"log_n": k}

    def fwd():
        L.check(lib.cqb_ntt_bn254_fr_dev(ctypes.c_void_p(a.data_ptr()), L.p64(d1.omega), k))

    def invf():
        L.check(lib.cqb_intt_bn254_fr_dev(ctypes.c_void_p(a.data_ptr()), L.p64(d1.omega_inv), L.p64(d1.ifft_divisor), k))

    for name, fn, nout, lg in (("forward", fwd, n, k), ("inverse", invf, n, k)):
        ms = timed(fn, reps)
        row[name] = {"ms": round(ms, 4), "gelem": round(nout / ms / 1e6, 3), "hbm_gbs": round(64 * nout / ms / 1e6, 1),
                     "hbm_frac": round(64 * nout / ms / 1e6 / HBM, 4),
                     "int_frac": round((nout / 2) * lg * 136 / (ms * 1e-3) / 1e12 / INT_PEAK, 3)}
    for name, dom in (("coset_n_to_2n", d1), ("coset_n_to_4n", d2)):
        if dom.extended_k > 28:
            continue
        en = 1 << dom.extended_k
        o = torch.empty(en * 32, dtype=torch.uint8, device="cuda")

        def cos():
            L.check(lib.cqb_coset_ntt_bn254_fr_dev(ctypes.c_void_p(a.data_ptr()), n, ctypes.c_void_p(o.data_ptr()), L.p64(dom.extended_omega),
                                                   dom.extended_k, L.p64(dom.g_coset), L.p64(dom.g_coset_inv)))

        ms = timed(cos, reps)
        row[name] = {"ms": round(ms, 4), "gelem_out": round(en / ms / 1e6, 3),
                     "int_frac": round((en / 2) * dom.extended_k * 136 / (ms * 1e-3) / 1e12 / INT_PEAK, 3)}
        if name == "coset_n_to_2n":
            def cinv():
                L.check(lib.cqb_coset_intt_bn254_fr_dev(ctypes.c_void_p(o.data_ptr()), dom.extended_k, L.p64(dom.extended_omega_inv),
                                                        L.p64(dom.extended_ifft_divisor), L.p64(dom.g_coset), L.p64(dom.g_coset_inv),
                                                        L.p64(dom.t_evaluations), dom.t_evaluations.shape[0]))

            ms = timed(cinv, reps)
            row["coset_inverse_2n_with_vanishing_division"] = {"ms": round(ms, 4), "gelem": round(en / ms / 1e6, 3)}
        del o
    res["ntt"].append(row)
    del a
    torch.cuda.empty_cache()
print(json.dumps(res))

"""tools/prove_workload.py — the synthetic CQ-prover-shaped workload of SURVEY.md §8(d) ("End-to-end prove ms").

No SHA-256 circuit exists in the reference, so "prove ms" is reported on the op list of §3.1 for a circuit with
A advice columns and one CQ static lookup over a table of N rows, at n = 2^k rows, through the HOST-pointer C ABI (the
calls the Rust prover would make): per proof
    A x commit_lagrange(n)                      plonk/prover.rs:356-360
    f commit_lagrange, m sparse MSM             static_lookup/prover.rs:165, 167-170
    A_cm, Q_A, A_0 sparse MSMs (|supp| = min(n, N))                    :245-257
    B iNTT(n), P = MSM(n-1) over the bound SRS slice, B_0 = commit(n)  :271, 299, 310
    f iNTT(n)                                                          :326-332
    random-poly commit(n)                        vanishing/prover.rs:58
    A x lagrange_to_coeff(n), (A + 2) x coeff_to_extended(n -> 2n)     plonk/prover.rs:587-603, evaluation.rs:317-334, 535-536
    quotient: divide_by_vanishing + extended_to_coeff(2n), 2 x commit(n)   vanishing/prover.rs:84-107
    1 GWC witness commit(n-1)                    gwc/prover.rs:80-86
Witness synthesis, evaluate_h's row program, transcript hashing and challenges are CPU work outside the hot path and are
NOT included. Returns milliseconds per proof (wall clock around the synchronous C-ABI calls).
"""
import ctypes
import os
import sys
import time

import numpy as np


def run(cq, k, A=8, table_log=16, reps=2, seed=0x70726F76, resident=False):
    """resident=False: every call takes HOST pointers (what the drop-in Rust shim does: each op copies its operands).
    resident=True : each column is uploaded ONCE per proof and stays in HBM (the *_dev entry points): commit_lagrange,
    lagrange_to_coeff and coeff_to_extended of a column share one upload, extended evaluations never leave the device
    (they feed evaluate_h, which is the next row of SURVEY §8f), only the 64 B commitments come back."""
    L = cq._lib
    lib = L.lib()
    n = 1 << k
    N = 1 << table_log
    Nt = max(N, n)
    dom = cq.EvaluationDomain(3, k)
    en = 1 << dom.extended_k

    def dev_bases(count, sd):
        d = ctypes.c_void_p()
        L.check(lib.cqb_dev_alloc(count * 64, ctypes.byref(d)))
        L.check(lib.cqb_synth_bases_dev(sd, 0, count, d))
        h = ctypes.c_uint64(0)
        L.check(lib.cqb_bases_register_device(d, count, ctypes.byref(h)))
        if count >= (1 << 16):
            L.check(lib.cqb_bases_precompute(h.value, 0))
        return d, h.value

    # SRS stand-ins (any distinct curve points give the same cost): g, g_lagrange, table g1 / lagrange / opening-at-0 / qs
    keep = [dev_bases(n, seed + i) for i in range(2)]
    (_, g), (_, g_lag) = keep
    tabs = [dev_bases(Nt, seed + 10 + i) for i in range(4)]
    t_g1, t_lag, t_op0, t_qs = (h for _, h in tabs)

    def pinned_scalars(count, sd):
        hp = ctypes.c_void_p()
        L.check(lib.cqb_host_alloc_pinned(count * 32, ctypes.byref(hp)))
        d = ctypes.c_void_p()
        L.check(lib.cqb_dev_alloc(count * 32, ctypes.byref(d)))
        L.check(lib.cqb_synth_scalars_dev(sd, 0, count, d))
        L.check(lib.cqb_memcpy_d2h(hp, d, count * 32))
        L.check(lib.cqb_dev_free(d))
        return hp

    cols = [pinned_scalars(n, seed + 100 + i) for i in range(A + 4)]  # advice, f, bs, random, h pieces reuse
    ext = pinned_scalars(en, seed + 200)
    ext_out = pinned_scalars(en, seed + 201)
    supp = min(n, N)
    idx = np.sort(np.random.default_rng(1).choice(Nt, supp, replace=False)).astype(np.uint32)
    sp = pinned_scalars(supp, seed + 300)
    out = np.zeros(8, np.uint64)
    inf = ctypes.c_int(0)
    u64 = L.u64p

    def msm(h, ptr, count, offset=0):
        L.check(lib.cqb_msm_bn254_g1(h, offset, ctypes.cast(ptr, u64), count, L.p64(out), ctypes.byref(inf)))

    def sparse(h):
        L.check(lib.cqb_msm_bn254_g1_sparse(h, idx.ctypes.data_as(L.u32p), ctypes.cast(sp, u64), supp, L.p64(out), ctypes.byref(inf)))

    def intt(ptr):
        L.check(lib.cqb_intt_bn254_fr(ctypes.cast(ptr, u64), L.p64(dom.omega_inv), L.p64(dom.ifft_divisor), k))

    def coset(ptr):
        L.check(lib.cqb_coset_ntt_bn254_fr(ctypes.cast(ptr, u64), n, ctypes.cast(ext_out, u64), L.p64(dom.extended_omega),
                                           dom.extended_k, L.p64(dom.g_coset), L.p64(dom.g_coset_inv)))

    dcols = []
    if resident:
        d_block = ctypes.c_void_p()   # the columns are contiguous so that their commitments can be batched
        L.check(lib.cqb_dev_alloc((A + 4) * n * 32, ctypes.byref(d_block)))
        dcols = [ctypes.c_void_p(d_block.value + i * n * 32) for i in range(A + 4)]
        d_ext, d_ext_out = ctypes.c_void_p(), ctypes.c_void_p()
        L.check(lib.cqb_dev_alloc(en * 32, ctypes.byref(d_ext)))
        L.check(lib.cqb_dev_alloc(en * 32, ctypes.byref(d_ext_out)))

    def msm_d(h, dptr, count, offset=0):
        L.check(lib.cqb_msm_bn254_g1_dev(h, offset, dptr, count, L.p64(out), ctypes.byref(inf)))

    def proof_resident():
        for i in range(A + 4):   # one upload per column per proof
            L.check(lib.cqb_memcpy_h2d(dcols[i], cols[i], n * 32))
        # the A advice commitments and f in ONE batched pass (plonk/prover.rs:356-360 + static_lookup/prover.rs:165)
        outs = np.zeros((A + 1, 8), np.uint64)
        infs = (ctypes.c_int * (A + 1))()
        L.check(lib.cqb_msm_bn254_g1_batch_dev(g_lag, 0, dcols[0], n, A + 1, L.p64(outs), infs))
        # permutation argument (plonk/permutation/prover.rs:46-200): grand product of every column set on the device, the
        # z commitments in one batched pass, lagrange_to_coeff; their coset NTTs feed evaluate_h below
        dw, last_z = 1, 1
        for st_ in range(nsets):
            cs_ = [dcols[a].value for a in range(3 * st_, min(3 * st_ + 3, A))]
            dw = product_set_dev(cs_, [d_perm_lag.value + a * n * 32 for a in range(3 * st_, 3 * st_ + len(cs_))], k, 3, 4, omega_int, dw, last_z,
                                 d_z.value + st_ * n * 32)
        zo = np.zeros((nsets, 8), np.uint64)
        zi = (ctypes.c_int * nsets)()
        L.check(lib.cqb_msm_bn254_g1_batch_dev(g_lag, 0, d_z, n, nsets, L.p64(zo), zi))
        for st_ in range(nsets):
            L.check(lib.cqb_intt_bn254_fr_dev(ctypes.c_void_p(d_z.value + st_ * n * 32), L.p64(dom.omega_inv), L.p64(dom.ifft_divisor), k))
        sparse(t_lag); sparse(t_lag); sparse(t_qs); sparse(t_op0)
        L.check(lib.cqb_intt_bn254_fr_dev(dcols[A + 1], L.p64(dom.omega_inv), L.p64(dom.ifft_divisor), k))
        msm_d(t_g1, dcols[A + 1], n - 1, offset=Nt - (n - 1))
        msm_d(g, dcols[A + 1], n)
        L.check(lib.cqb_intt_bn254_fr_dev(dcols[A], L.p64(dom.omega_inv), L.p64(dom.ifft_divisor), k))
        msm_d(g, dcols[A + 2], n)
        for a in range(A):
            L.check(lib.cqb_intt_bn254_fr_dev(dcols[a], L.p64(dom.omega_inv), L.p64(dom.ifft_divisor), k))
        evaluate_h_resident()   # the (A + 2) coset NTTs + custom gates + permutation + CQ terms, device-resident
        L.check(lib.cqb_coset_intt_bn254_fr_dev(d_ext, dom.extended_k, L.p64(dom.extended_omega_inv), L.p64(dom.extended_ifft_divisor),
                                                L.p64(dom.g_coset), L.p64(dom.g_coset_inv), L.p64(dom.t_evaluations),
                                                dom.t_evaluations.shape[0]))
        msm_d(g, d_ext, n); msm_d(g, ctypes.c_void_p(d_ext.value + n * 32), n)   # the two h pieces, straight from the device
        msm_d(g, dcols[A + 3], n - 1)

    # ---- the row program of evaluate_h (custom gates + permutation + CQ term) for the resident mode -------------------
    if resident:
        import random

        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from tests.evalh_common import random_expr
        from sha2_on_cq_halo2_b200.evaluation import Expr, cq_lookup_h_dev, custom_gates_evaluator, permutation_h_dev

        rng = random.Random(11)
        gate_polys = [random_expr(rng, Expr, 5, ncols=(2, A, 1), nchal=1) for _ in range(12)]   # 12 gate polynomials
        ev = custom_gates_evaluator(gate_polys)
        rot_scale = 1 << (dom.extended_k - k)
        d_ext_cols = ctypes.c_void_p()   # coset evaluations of the A advice columns + f + b + 2 fixed + 1 instance + A permutation cosets
        n_ext_cols = A + 2 + 3 + A + 3 + 2
        L.check(lib.cqb_dev_alloc(n_ext_cols * en * 32, ctypes.byref(d_ext_cols)))
        ext_ptr = [d_ext_cols.value + i * en * 32 for i in range(n_ext_cols)]
        for i in range(A + 2, n_ext_cols):   # key material (fixed / permutation cosets, l0 / l_last / l_active, z columns): resident, filled once
            L.check(lib.cqb_synth_scalars_dev(seed + 500 + i, 0, en, ctypes.c_void_p(ext_ptr[i])))
        chal1 = np.zeros((1, 4), np.uint64)
        bgty = [np.array([3 + i, 0, 0, 0], np.uint64) for i in range(4)]
        nsets = (A + 2) // 3   # chunk_len = cs.degree() - 2 = 3 columns per permutation set
        from sha2_on_cq_halo2_b200.permutation import product_set_dev
        from sha2_on_cq_halo2_b200.fields import fr_from_limbs
        omega_int = fr_from_limbs(dom.omega)
        d_perm_lag, d_z = ctypes.c_void_p(), ctypes.c_void_p()   # pkey.permutations (Lagrange, key material) and the z vectors
        L.check(lib.cqb_dev_alloc(A * n * 32, ctypes.byref(d_perm_lag)))
        L.check(lib.cqb_dev_alloc(nsets * n * 32, ctypes.byref(d_z)))
        L.check(lib.cqb_synth_scalars_dev(seed + 900, 0, A * n, d_perm_lag))

    def evaluate_h_resident():
        """plonk/prover.rs:606-624: coset NTT of every advice / CQ polynomial, then the row program, all in HBM"""
        for a in list(range(A)) + [A, A + 1]:
            L.check(lib.cqb_coset_ntt_bn254_fr_dev(dcols[a], n, ctypes.c_void_p(ext_ptr[a]), L.p64(dom.extended_omega), dom.extended_k,
                                                   L.p64(dom.g_coset), L.p64(dom.g_coset_inv)))
        for st_ in range(min(nsets, 2)):   # the z polynomials' cosets (evaluation.rs:376-452 reads them)
            L.check(lib.cqb_coset_ntt_bn254_fr_dev(ctypes.c_void_p(d_z.value + st_ * n * 32), n, ctypes.c_void_p(ext_ptr[2 * A + 8 + st_]),
                                                   L.p64(dom.extended_omega), dom.extended_k, L.p64(dom.g_coset), L.p64(dom.g_coset_inv)))
        L.check(lib.cqb_memcpy_d2d(d_ext, ctypes.c_void_p(ext_ptr[n_ext_cols - 1]), en * 32))   # values := 0-like start (any vector)
        fixed = ext_ptr[A + 2:A + 4]
        inst = ext_ptr[A + 4:A + 5]
        ev.evaluate_dev(fixed, ext_ptr[:A], inst, chal1, bgty[0], bgty[1], bgty[2], bgty[3], d_ext.value, en, rot_scale)
        perm = ext_ptr[A + 5:A + 5 + A]
        l0, l_last, l_act = ext_ptr[2 * A + 5:2 * A + 8]
        sets = ext_ptr[2 * A + 8:2 * A + 10][:max(1, min(nsets, 2))]
        ncols_p = min(A, 3 * len(sets))
        permutation_h_dev(d_ext.value, en, rot_scale, -(5 + 1), 3, sets, ext_ptr[:ncols_p], perm[:ncols_p], l0, l_last, l_act, bgty[0], bgty[1],
                          bgty[3], dom.extended_omega)
        cq_lookup_h_dev(d_ext.value, ext_ptr[A + 1], ext_ptr[A], l_act, bgty[0], bgty[3], en)

    def proof_host():
        for a in range(A):
            msm(g_lag, cols[a], n)
        msm(g_lag, cols[A], n)          # f
        sparse(t_lag)                   # m
        sparse(t_lag); sparse(t_qs); sparse(t_op0)   # A, Q_A, A_0
        intt(cols[A + 1])               # B
        msm(t_g1, cols[A + 1], n - 1, offset=Nt - (n - 1))   # P over the degree-bound slice
        msm(g, cols[A + 1], n)          # B_0
        intt(cols[A])                   # f -> coeff
        msm(g, cols[A + 2], n)          # random poly
        for a in range(A):
            intt(cols[a])
        for a in range(A):
            coset(cols[a])
        coset(cols[A]); coset(cols[A + 1])
        L.check(lib.cqb_coset_intt_bn254_fr(ctypes.cast(ext, u64), dom.extended_k, L.p64(dom.extended_omega_inv),
                                            L.p64(dom.extended_ifft_divisor), L.p64(dom.g_coset), L.p64(dom.g_coset_inv),
                                            L.p64(dom.t_evaluations), dom.t_evaluations.shape[0]))
        msm(g, cols[A + 3], n); msm(g, cols[A + 2], n)   # two h pieces
        msm(g, cols[A + 3], n - 1)      # GWC witness

    proof = proof_resident if resident else proof_host
    proof()  # warm-up (twiddle tables, scratch growth)
    L.check(lib.cqb_sync())
    l0 = lib.cqb_launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        proof()
    L.check(lib.cqb_sync())
    ms = (time.perf_counter() - t0) * 1e3 / reps
    launches = (lib.cqb_launch_count() - l0) // reps
    for p in cols + [ext, ext_out, sp]:
        L.check(lib.cqb_host_free_pinned(p))
    if resident:
        for d in [d_block, d_ext, d_ext_out, d_ext_cols, d_perm_lag, d_z]:
            L.check(lib.cqb_dev_free(d))
    for d, h in keep + tabs:
        L.check(lib.cqb_bases_free(h))
        L.check(lib.cqb_dev_free(d))
    n_msm = A + 1 + 4 + 2 + 1 + 2 + 1
    return {"k": k, "mode": "device-resident polynomials, advice commitments batched, permutation grand products + z commitments, evaluate_h "
                            "row program (12 gate polynomials, permutation + CQ terms) on the device" if resident else "host-pointer calls (drop-in)", "advice_columns": A, "table_rows": N, "ms_per_proof": ms, "gpu_launches_per_proof": int(launches),
            "ops": {"dense_msm": n_msm - 4, "sparse_msm": 4, "intt_n": A + 2, "coset_ntt_2n": A + 2, "coset_intt_2n": 1}}


if __name__ == "__main__":
    import json
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import cqb200

    cqb200._lib.init(0)
    for k in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["16", "20"])]:
        print(json.dumps(run(cqb200, k)), flush=True)
        print(json.dumps(run(cqb200, k, resident=True)), flush=True)

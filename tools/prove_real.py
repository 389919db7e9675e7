"""tools/prove_real.py — "SHA2-CQ prove ms" as a REAL proof instead of an op list: a complete create_proof
(sha2_on_cq_halo2_b200.prover, reference plonk/prover.rs:37-797 + gwc/prover.rs:42-86) of a CQ-lookup circuit of the shape the sha
crate's tables call for (the reference has no SHA circuit; sha/src/tables.rs only generates 2^16-row spread tables):
A advice columns, one static lookup of the tuple (a0, a1) in two 2^logN-row tables, a permutation argument over all columns,
KZG / GWC, Blake2b transcript (hashlib; transcript.rs:199-240). Witness synthesis, the m_sparse map and the transcript hash are CPU
work outside the path, as in the reference; the timed region is create_proof itself: witness upload, every commitment, NTT,
evaluate_h, evaluation and opening — and each proof is CHECKED: the quotient identity at x (prover.expected_h_eval, what
plonk/verifier.rs computes from the evaluations in the proof) must hold.

Not test infrastructure: nothing here touches oracle/."""
import hashlib
import time

import numpy as np

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
_RINV_Q = pow(1 << 256, -1, Q_MOD)


class Blake2bTranscript:
    """halo2_proofs/src/transcript.rs:199-240 (Blake2bWrite) and :297-315 (Challenge255)"""

    def __init__(self):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.proof = bytearray()

    def common_scalar(self, v):
        self.state.update(b"\x02" + int(v).to_bytes(32, "little"))

    def write_scalar(self, v):
        self.common_scalar(v)
        self.proof += int(v).to_bytes(32, "little")

    def write_point(self, affine_limbs):
        a = np.asarray(affine_limbs, dtype=np.uint64)
        x = sum(int(a[i]) << (64 * i) for i in range(4)) * _RINV_Q % Q_MOD
        y = sum(int(a[4 + i]) << (64 * i) for i in range(4)) * _RINV_Q % Q_MOD
        self.state.update(b"\x01" + x.to_bytes(32, "little") + y.to_bytes(32, "little"))
        b = bytearray(x.to_bytes(32, "little"))
        b[31] |= (y & 1) << 7                      # derive/curve.rs:635-646
        self.proof += b

    def squeeze_challenge_scalar(self):
        self.state.update(b"\x00")
        return int.from_bytes(self.state.copy().digest(), "little") % R_MOD


def _fr_array(ints):
    """canonical ints -> (n, 4) uint64 Montgomery limbs"""
    from sha2_on_cq_halo2_b200.fields import fr_to_limbs

    return np.stack([fr_to_limbs(v) for v in ints])


def _fr_array_fast(ints):
    """the same for many small values: vectorised through Python big ints in one pass"""
    Rm = (1 << 256) % R_MOD
    out = np.empty((len(ints), 4), np.uint64)
    mask = (1 << 64) - 1
    for i, v in enumerate(ints):
        m = v * Rm % R_MOD
        out[i, 0], out[i, 1], out[i, 2], out[i, 3] = m & mask, (m >> 64) & mask, (m >> 128) & mask, m >> 192
    return out


def circuit_arrays(k, log_table, n_advice, seed=1):
    """the circuit as plain host arrays (no device, no library): toxic s, table values, advice witness, sigma columns, the lookup's
    m_sparse in key order, the rng draws"""
    from sha2_on_cq_halo2_b200.fields import FR_ROOT_OF_UNITY, FR_S, fr_to_limbs
    from sha2_on_cq_halo2_b200.permutation import FR_DELTA

    n, N, A = 1 << k, 1 << log_table, n_advice
    cs_degree, bf = 4, 5
    usable = n - (bf + 1)
    rng = np.random.default_rng(seed)
    s = fr_to_limbs(int(rng.integers(1, 1 << 62)) * 0x9E3779B97F4A7C15 % R_MOD)
    # spread-table-like values: distinct 30-bit values, second table offset (sha/src/tables.rs generates (x, f(x)) tuples)
    t0 = rng.choice(1 << 30, N, replace=False).astype(np.int64)
    tvals = [[int(v) for v in t0], [int(v) + (1 << 40) for v in rng.permutation(t0)]]
    rows = rng.integers(0, N, usable)
    adv = [[tv[r] for r in rows] + [int(v) for v in rng.integers(0, 1 << 50, n - usable)] for tv in tvals]
    for j in range(2, A):
        adv.append(adv[j % 2][:usable] + [int(v) for v in rng.integers(0, 1 << 50, n - usable)])
    w = FR_ROOT_OF_UNITY
    for _ in range(k, FR_S):
        w = w * w % R_MOD
    wp = [1] * n
    for i in range(1, n):
        wp[i] = wp[i - 1] * w % R_MOD
    sig = [[pow(FR_DELTA, j, R_MOD) * wp[i] % R_MOD for i in range(n)] for j in range(A)]
    for j in range(2, A):                         # copy constraints between column j and column j % 2 (equal cells)
        for i in range(j, usable, 4):
            sig[j][i], sig[j % 2][i] = sig[j % 2][i], sig[j][i]
    m = {}
    for r in rows:
        m[int(r)] = m.get(int(r), 0) + 1
    idx = np.array(sorted(m), dtype=np.uint32)
    nsets = (A + cs_degree - 3) // (cs_degree - 2)
    return {"k": k, "N": N, "A": A, "cs_degree": cs_degree, "bf": bf, "s": s, "vk_repr": 0xC0DE + k,
            "tables": [_fr_array_fast(v) for v in tvals], "advice": [_fr_array_fast(c) for c in adv], "sigma": [_fr_array_fast(sg) for sg in sig],
            "idx": idx, "mult": _fr_array_fast([m[int(i)] for i in idx]),
            "permutation_blinds": [_fr_array_fast([int(v) for v in rng.integers(1, 1 << 62, bf)]) for _ in range(nsets)],
            "random_poly": _fr_array_fast([int(v) for v in rng.integers(1, 1 << 62, n)])}


def build_circuit(cq, k, log_table, n_advice, seed=1):
    """proving key + witness of the lookup circuit; everything that keygen would prepare (SRS, tables with their cached quotients,
    sigma polynomials) is resident on the device when this returns"""
    from sha2_on_cq_halo2_b200 import prover as PR

    c = circuit_arrays(k, log_table, n_advice, seed)
    n, N, A = 1 << k, c["N"], c["A"]
    params = cq.ParamsKZG.setup_from_toxic_waste(k, c["s"])
    Nt = max(N, n)
    tsrs = cq.TableSRS.setup_from_toxic_waste(N - 1, c["s"])
    big = tsrs if Nt == N else cq.TableSRS.setup_from_toxic_waste(Nt - 1, c["s"], precompute=False)
    # b0_g1_bound = the last n - 1 powers of the length-Nt table SRS (my_test.rs:205, static_lookup.rs:149)
    bound_host = big.g1.to_host()[Nt - (n - 1):]
    bound = cq.DeviceBases(np.ascontiguousarray(bound_host))
    tables = [cq.cq.StaticTableValues(v, tsrs.g1) for v in c["tables"]]
    lk = PR.StaticLookup([0, 1], tsrs, tables, bound)
    pk = PR.ProvingKey(params, k, c["cs_degree"], c["bf"], list(range(A)), c["sigma"], [(j, 0) for j in range(A)], [lk],
                       vk_transcript_repr=c["vk_repr"])
    rnd = {"permutation_blinds": c["permutation_blinds"], "random_poly": c["random_poly"]}
    keep = (params, tsrs, big, bound, tables)
    return pk, c["advice"], (c["idx"], c["mult"]), rnd, keep


def run(cq, k, log_table=16, n_advice=8, reps=3):
    from sha2_on_cq_halo2_b200 import prover as PR

    L = cq._lib
    lib = L.lib()
    pk, witness, m_sparse, rnd, keep = build_circuit(cq, k, min(log_table, 16), n_advice)
    ok, proof_len, launches = True, 0, 0
    PR.create_proof(pk, witness, [m_sparse], rnd, Blake2bTranscript())  # warm-up: twiddle tables, scratch growth
    L.check(lib.cqb_sync())
    times = []
    for _ in range(reps):
        t = Blake2bTranscript()
        l0 = lib.cqb_launch_count()
        t0 = time.perf_counter()
        info = PR.create_proof(pk, witness, [m_sparse], rnd, t)
        L.check(lib.cqb_sync())
        times.append((time.perf_counter() - t0) * 1e3)
        launches = lib.cqb_launch_count() - l0
        ok = ok and PR.expected_h_eval(pk, info) == info["h_eval"]
        proof_len = len(t.proof)
    pk.free()
    params, tsrs, big, bound, tables = keep
    for tb in tables:
        tb.free()
    bound.free()
    if big is not tsrs:
        big.free()
    tsrs.free()
    params.free()
    return {"k": k, "table_rows": 1 << min(log_table, 16), "advice_columns": n_advice, "ms_per_proof": float(np.median(times)),
            "ms_min": float(min(times)), "proof_bytes": proof_len, "quotient_identity_holds": bool(ok), "gpu_launches_per_proof": int(launches),
            "h2d_bytes_per_proof": n_advice * (1 << k) * 32 + (1 << k) * 32}


if __name__ == "__main__":
    import json
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import cqb200

    cqb200._lib.init(0)
    for kk in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "12,16").split(",")]:
        print(json.dumps(run(cqb200, kk)), flush=True)

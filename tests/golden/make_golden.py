"""tests/golden/make_golden.py — regenerates tests/golden/vectors.json.

The reference (Rust) cannot be built or run in this image and holds no golden vectors for this path, so
these fixtures are produced by the CPU oracle (oracle/bn254_oracle.c, the line-by-line restatement of the reference) and
every MSM / NTT value is cross-checked here against the independent Python big-integer model (oracle/pyref.py) before it
is written. They pin the oracle and the CUDA path against silent drift; they are not outputs of the reference binary.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import oracle_lib as O
from oracle import pyref as P


def hx(a):
    return [f"{int(v):016x}" for v in np.asarray(a, dtype=np.uint64).reshape(-1)]


out = {"about": "oracle-generated, pyref-cross-checked; Montgomery-form little-endian u64 limbs as hex", "msm": [], "ntt": [], "kzg": [], "cq": []}
for n, seed in ((1, 11), (5, 12), (33, 13), (300, 14)):
    sc = O.synth_scalars(seed, n)
    bs = O.synth_bases(seed + 100, n, 2)
    if n >= 5:
        sc[0] = 0
        bs[3] = 0
    _, aff = O.best_multiexp(sc, bs, 3)
    exp = P.msm(P.fr_array_to_ints(sc), P.g1_affine_to_ints(bs))
    assert P.g1_affine_to_ints(aff)[0] == exp
    out["msm"].append({"n": n, "scalar_seed": seed, "base_seed": seed + 100, "zeroed_scalar": 0 if n >= 5 else None,
                       "identity_base": 3 if n >= 5 else None, "affine": hx(aff), "compressed": O.g1_to_bytes(aff).hex()})
for k, seed in ((0, 21), (1, 22), (4, 23), (9, 24)):
    a = O.synth_scalars(seed, 1 << k)
    w = P.omega_for(k)
    res = O.best_fft(a, P.int_to_limbs(P.to_mont(w, P.R_MOD)), k, 2)
    assert P.fr_array_to_ints(res) == P.dft(P.fr_array_to_ints(a), w)
    d = O.domain_new(3, k) if k >= 1 else None
    row = {"log_n": k, "seed": seed, "first": hx(res[0]), "last": hx(res[-1]), "xor_of_all_limbs": hx(np.bitwise_xor.reduce(res, axis=0))}
    if d is not None:
        ext = O.coeff_to_extended(d, a)
        row["coset_extended_k"] = int(d.extended_k)
        row["coset_xor"] = hx(np.bitwise_xor.reduce(ext, axis=0))
        row["quotient_xor"] = hx(np.bitwise_xor.reduce(O.extended_to_coeff(d, O.divide_by_vanishing_poly(d, ext)), axis=0))
    out["ntt"].append(row)
s = O.synth_scalars(31, 1)[0]
g, gl = O.params_setup(5, s)
a = P.fr_array_from_ints(list(range(32)))
_, c = O.best_multiexp(a, gl, 2)
out["kzg"].append({"k": 5, "toxic_seed": 31, "g_last": hx(g[-1]), "g_lagrange_last": hx(gl[-1]), "commit_lagrange_0_to_31": hx(c)})
g1, l1, op0 = O.table_srs_setup(16, s)
vals = O.synth_scalars(32, 16)
qs = O.cq_table_qs(vals, g1, 2)
out["cq"].append({"N": 16, "toxic_seed": 31, "value_seed": 32, "opening_at_0_last": hx(op0[-1]), "qs_first": hx(qs[0]), "qs_last": hx(qs[-1])})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json"), "w"), indent=1)
print("wrote vectors.json")

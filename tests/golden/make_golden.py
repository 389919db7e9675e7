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
# grand products (permutation/prover.rs:82-166, lookup/prover.rs:173-262), cross-checked against Python integers
DELTA = 0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2
out["products"] = []
k, ncols = 5, 3
n = 1 << k
cols = [O.synth_scalars(41 + j, n) for j in range(ncols)]
perms = [O.synth_scalars(51 + j, n) for j in range(ncols)]
beta, gamma, last_z = (P.fr_array_to_ints(O.synth_scalars(61 + j, 1))[0] for j in range(3))
one = lambda x: P.fr_array_from_ints([x])[0]  # noqa: E731
w = P.omega_for(k)
z, dw = O.permutation_product(cols, perms, one(beta), one(gamma), one(w), one(1), one(last_z))
ci, pi = [P.fr_array_to_ints(c) for c in cols], [P.fr_array_to_ints(p_) for p_ in perms]
mv = [1] * n
for j in range(ncols):
    for i in range(n):
        mv[i] = mv[i] * (beta * pi[j][i] + gamma + ci[j][i]) % P.R_MOD
mv = [pow(v, -1, P.R_MOD) for v in mv]
d = 1
for j in range(ncols):
    cur = d
    for i in range(n):
        mv[i] = mv[i] * (cur * beta + gamma + ci[j][i]) % P.R_MOD
        cur = cur * w % P.R_MOD
    d = d * DELTA % P.R_MOD
zz = [last_z]
for i in range(1, n):
    zz.append(zz[-1] * mv[i - 1] % P.R_MOD)
assert P.fr_array_to_ints(z) == zz and P.fr_array_to_ints(dw[None, :])[0] == d
out["products"].append({"kind": "permutation", "k": k, "ncols": ncols, "col_seed": 41, "perm_seed": 51, "challenge_seed": 61, "z_last": hx(z[-1]),
                        "z_xor": hx(np.bitwise_xor.reduce(z, axis=0)), "deltaomega_out": hx(dw)})
a, sv, ap, sp = (O.synth_scalars(71 + j, n) for j in range(4))
lz = O.lookup_product(a, sv, ap, sp, one(beta), one(gamma))
ai, si, api, spi = (P.fr_array_to_ints(x) for x in (a, sv, ap, sp))
acc, lzz = 1, [1]
for i in range(n - 1):
    acc = acc * pow((beta + api[i]) * (gamma + spi[i]) % P.R_MOD, -1, P.R_MOD) * (ai[i] + beta) % P.R_MOD * (si[i] + gamma) % P.R_MOD
    lzz.append(acc)
assert P.fr_array_to_ints(lz) == lzz
out["products"].append({"kind": "lookup", "k": k, "seed": 71, "challenge_seed": 61, "z_last": hx(lz[-1]), "z_xor": hx(np.bitwise_xor.reduce(lz, axis=0))})
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "vectors.json"), "w"), indent=1)
print("wrote vectors.json")

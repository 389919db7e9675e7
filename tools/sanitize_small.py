"""tools/sanitize_small.py — one tiny invocation of every kernel family, for `compute-sanitizer --tool memcheck`."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import cqb200
from oracle import oracle_lib as O
from oracle import pyref as P

cqb200._lib.init(0)
L, lib = cqb200._lib, cqb200._lib.lib()
for n in (1, 7, 300, 5000):
    sc, bs = O.synth_scalars(1, n), O.synth_bases(2, n, 2)
    assert np.array_equal(cqb200.best_multiexp(sc, bs).to_affine(), O.best_multiexp(sc, bs, 2)[1])
n = 1 << 13
sc, bs = O.synth_scalars(3, n), O.synth_bases(4, n, 2)
dev = cqb200.DeviceBases(bs, precompute=True, window_bits=12)
exp = O.best_multiexp(sc, bs, 2)[1]
assert np.array_equal(dev.msm(sc).to_affine(), exp)
sk = sc.copy(); sk[:] = sk[0]
assert np.array_equal(dev.msm(sk).to_affine(), O.best_multiexp(sk, bs, 2)[1])
idx = np.arange(0, n, 3, dtype=np.uint32)
assert np.array_equal(dev.msm_sparse(idx, sc[: idx.shape[0]]).to_affine(), O.sparse_commit(bs, idx, sc[: idx.shape[0]]))
dev.free()
for k in (1, 5, 9, 12):
    a = O.synth_scalars(5, 1 << k)
    w = P.int_to_limbs(P.to_mont(P.omega_for(k), P.R_MOD))
    e = O.best_fft(a, w, k, 2)
    cqb200.best_fft(a, w, k)
    assert np.array_equal(a, e)
d = cqb200.EvaluationDomain(3, 7)
od = O.domain_new(3, 7)
c = O.synth_scalars(6, 128)
ext = d.coeff_to_extended(c)
assert np.array_equal(ext.values, O.coeff_to_extended(od, c))
assert np.array_equal(d.extended_to_coeff(d.divide_by_vanishing_poly(ext)), O.extended_to_coeff(od, O.divide_by_vanishing_poly(od, ext.values)))
s = O.synth_scalars(7, 1)[0]
p = cqb200.ParamsKZG.setup_from_toxic_waste(6, s, precompute=False)
g, gl = O.params_setup(6, s)
assert np.array_equal(p.g_lagrange.to_host(), gl)
p.downsize(4)
assert np.array_equal(p.g_lagrange.to_host(), O.params_setup(4, s)[1])
p.free()
t = cqb200.TableSRS.setup_from_toxic_waste(15, s, precompute=False)
vals = O.synth_scalars(8, 16)
tv = cqb200.cq.StaticTableValues(vals, t.g1)
assert np.array_equal(tv.qs.to_host(), O.cq_table_qs(vals, t.g1.to_host(), 2))
tv.free(); t.free()
a = O.synth_scalars(9, 1000)
x = O.synth_scalars(10, 1)[0]
assert np.array_equal(cqb200.eval_polynomial(a, x), O.eval_polynomial(a, x))
assert np.array_equal(cqb200.kate_division(a, x), O.kate_division(a, x))
# ---- round 2 kernels: batched-affine accumulation (+ mark / merge), batch_normalize, G2, axpy / batched evaluation, a complete proof
for seg in (3, 5):
    L.check(lib.cqb_msm_set_accumulator(2, seg))
    n = 3000
    sc, bs = O.synth_scalars(20 + seg, n), O.synth_bases(21 + seg, n, 2)
    bs[5] = 0
    bs[7] = bs[6]; sc[7] = sc[6]
    assert np.array_equal(cqb200.best_multiexp(sc, bs).to_affine(), O.best_multiexp(sc, bs, 2)[1])
    dev = cqb200.DeviceBases(bs, precompute=True, window_bits=11)
    assert np.array_equal(dev.msm(sc).to_affine(), O.best_multiexp(sc, bs, 2)[1])
    sk = sc.copy(); sk[:] = sk[0]
    assert np.array_equal(dev.msm(sk).to_affine(), O.best_multiexp(sk, bs, 2)[1])
    dev.free()
L.check(lib.cqb_msm_set_accumulator(0, 0))
jac = np.stack([O.g1_mul_a(b, O.synth_scalars(30 + i, 1)[0]) for i, b in enumerate(O.synth_bases(31, 40, 1))])
jac[3] = 0
assert np.array_equal(cqb200.batch_normalize(jac), O.g1_batch_normalize(jac))
msm = cqb200.MSMKZG()
scj = O.synth_scalars(32, 40)
for i in range(40):
    msm.append_term(scj[i], jac[i])
assert np.array_equal(msm.eval().to_affine(), O.best_multiexp(scj, O.g1_batch_normalize(jac), 2)[1])
g2 = cqb200.kzg.g2_powers(s, 9)
assert np.array_equal(g2, O.g2_powers(s, 9))
o16, inf = np.zeros(16, np.uint64), ctypes.c_int(0)
sc2 = O.synth_scalars(33, 9)
L.check(lib.cqb_msm_bn254_g2(L.p64(g2), L.p64(sc2), 9, L.p64(o16), ctypes.byref(inf)))
assert np.array_equal(o16, O.g2_msm(g2, sc2))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
import prove_real
from sha2_on_cq_halo2_b200 import prover as PR

pk, witness, m_sparse, rnd, keep = prove_real.build_circuit(cqb200, 6, 5, 3)
info = PR.create_proof(pk, witness, [m_sparse], rnd, prove_real.Blake2bTranscript())
assert PR.expected_h_eval(pk, info) == info["h_eval"]
pk.free()
print("sanitize_small: all ok; launches", lib.cqb_launch_count())

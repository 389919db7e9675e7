import cProfile, pstats, sys, os, io
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import cqb200, prove_real
from sha2_on_cq_halo2_b200 import prover as PR
cqb200._lib.init(0)
for k in (14, 16):
    pk, witness, m_sparse, rnd, keep = prove_real.build_circuit(cqb200, k, 16, 8)
    PR.create_proof(pk, witness, [m_sparse], rnd, prove_real.Blake2bTranscript())
    pr = cProfile.Profile(); pr.enable()
    for _ in range(3):
        PR.create_proof(pk, witness, [m_sparse], rnd, prove_real.Blake2bTranscript())
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(14); print('k', k); print(s.getvalue()[:3500])

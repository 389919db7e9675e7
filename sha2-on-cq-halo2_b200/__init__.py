"""sha2-on-cq-halo2_b200 — B200-native (sm_100a) BN254 MSM + Fr NTT backend for the halo2 SHA2-on-CQ prover's hot path.

Layout: csrc/ (hand-written CUDA kernels + the C ABI of include/cqb200.h, built into libcqb200.so) and the host-side
mirror of the reference interface for this path: arithmetic (best_multiexp, best_fft), domain (EvaluationDomain),
kzg (ParamsKZG.commit / commit_lagrange, TableSRS), cq (CQ prover commit calls), permutation (grand products), sharded (point-range-sharded MSM over
the GPUs of one box). Import name: `sha2_on_cq_halo2_b200` via the repo-root shim `cqb200.py`.
"""
from . import _lib  # noqa: F401
from . import arithmetic, cq, domain, evaluation, fields, kzg, lookup, permutation, prover  # noqa: F401
from .arithmetic import G1, best_fft, best_multiexp, eval_polynomial, kate_division  # noqa: F401
from .domain import EvaluationDomain  # noqa: F401
from .kzg import MSMKZG, DeviceBases, ParamsKZG, TableSRS, batch_normalize  # noqa: F401

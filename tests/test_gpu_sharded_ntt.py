"""Building blocks of the distributed four-step NTT on one GPU: batched transforms, the omega-power twiddle step, the 32-byte
element transpose — each against the oracle — and ShardedNTT with the CUDA backend at world = 1 (all three transposes, both
batched transforms and the twiddles run; only the all-to-all is trivial) against the single-call transform. The multi-rank
exchange logic is covered on CPU (tests/test_sharded_cpu.py, gloo) and on 2-8 GPUs by tools/bench_sharded_ntt.py."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyref as P  # noqa: E402


@pytest.fixture(scope="module")
def env():
    import torch

    import cqb200

    cqb200._lib.init(0)
    return cqb200, torch


def _to_dev(torch, arr):
    return torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).cuda()


def _to_host(t):
    return t.cpu().numpy().view(np.uint64).reshape(-1, 4)


@pytest.mark.parametrize("log_n,batch", [(1, 5), (4, 3), (9, 17), (13, 4)])
def test_batched_ntt_matches_oracle_per_member(env, oracle, log_n, batch):
    cq, torch = env
    L, lib = cq._lib, cq._lib.lib()
    n = 1 << log_n
    a = oracle.synth_scalars(0x600 + log_n, n * batch)
    w = P.int_to_limbs(P.to_mont(P.omega_for(log_n), P.R_MOD))
    exp = np.concatenate([oracle.best_fft(np.ascontiguousarray(a[b * n:(b + 1) * n]), w, log_n, 1) for b in range(batch)])
    t = _to_dev(torch, a)
    L.check(lib.cqb_ntt_bn254_fr_batch_dev(ctypes.c_void_p(t.data_ptr()), L.p64(w), log_n, batch))
    L.check(lib.cqb_sync())
    assert np.array_equal(_to_host(t), exp)


@pytest.mark.parametrize("rows,cols", [(1, 1), (3, 70), (32, 32), (33, 65), (256, 1024)])
def test_transpose_and_omega_powers(env, oracle, rows, cols):
    cq, torch = env
    L, lib = cq._lib, cq._lib.lib()
    a = oracle.synth_scalars(0x700 + rows, rows * cols)
    t = _to_dev(torch, a)
    out = torch.empty_like(t)
    L.check(lib.cqb_fr_transpose_dev(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(out.data_ptr()), rows, cols))
    L.check(lib.cqb_sync())
    assert np.array_equal(_to_host(out).reshape(cols, rows, 4), a.reshape(rows, cols, 4).transpose(1, 0, 2))
    if rows * cols <= 4096:
        log_n, row0 = 12, 37
        w_int = P.omega_for(log_n)
        L.check(lib.cqb_fr_mul_omega_powers_dev(ctypes.c_void_p(t.data_ptr()), rows, cols, row0, L.p64(P.int_to_limbs(P.to_mont(w_int, P.R_MOD))), log_n))
        L.check(lib.cqb_sync())
        vals = P.fr_array_to_ints(a)
        exp = [v * pow(w_int, (row0 + i // cols) * (i % cols), P.R_MOD) % P.R_MOD for i, v in enumerate(vals)]
        assert P.fr_array_to_ints(_to_host(t)) == exp


@pytest.mark.parametrize("log_n", [2, 7, 12, 17])
def test_sharded_ntt_world1_equals_single_call(env, oracle, log_n):
    cq, torch = env
    from sha2_on_cq_halo2_b200.sharded import CudaNttBackend, ShardedNTT

    L, lib = cq._lib, cq._lib.lib()
    n = 1 << log_n
    a = oracle.synth_scalars(0x800 + log_n, n)
    sn = ShardedNTT(CudaNttBackend("cuda:0"), log_n)
    got = sn.forward(_to_dev(torch, a))
    L.check(lib.cqb_sync())
    w = P.int_to_limbs(P.to_mont(sn.omega, P.R_MOD))
    if log_n <= 12:
        exp = oracle.best_fft(a.copy(), w, log_n, 2)
    else:
        ref = _to_dev(torch, a)
        L.check(lib.cqb_ntt_bn254_fr_dev(ctypes.c_void_p(ref.data_ptr()), L.p64(w), log_n))
        L.check(lib.cqb_sync())
        exp = _to_host(ref)
    assert np.array_equal(_to_host(got), exp)
    back = sn.inverse(got)
    L.check(lib.cqb_sync())
    assert np.array_equal(_to_host(back), a)


@pytest.mark.parametrize("log_n,batch,seg_log", [(6, 8, 2), (10, 4, 0), (12, 16, 5), (9, 3, 9)])
def test_mapped_batched_ntt(env, oracle, log_n, batch, seg_log):
    """cqb_ntt_bn254_fr_batch_map_dev: segmented gather, transposed store and fused omega-power twiddle against the plain
    batched transform + explicit permutations on the host"""
    cq, torch = env
    L, lib = cq._lib, cq._lib.lib()
    n, seg = 1 << log_n, 1 << seg_log
    nat = oracle.synth_scalars(0x900 + log_n, n * batch).reshape(batch, n, 4)            # member-major natural layout
    src = np.ascontiguousarray(nat.reshape(batch, n // seg, seg, 4).transpose(1, 0, 2, 3))  # [segment index][member][segment]
    w_int = P.omega_for(log_n)
    w = P.int_to_limbs(P.to_mont(w_int, P.R_MOD))
    big_log, row0 = 14, 5
    wb_int = P.omega_for(big_log)
    wb = P.int_to_limbs(P.to_mont(wb_int, P.R_MOD))
    exp = np.stack([oracle.best_fft(np.ascontiguousarray(nat[b]), w, log_n, 1) for b in range(batch)])
    for b in range(batch):
        vals = P.fr_array_to_ints(exp[b])
        exp[b] = P.fr_array_from_ints([v * pow(wb_int, (row0 + b) * i, P.R_MOD) % P.R_MOD for i, v in enumerate(vals)])
    exp_t = np.ascontiguousarray(exp.transpose(1, 0, 2))                                    # [idx][member]
    d_src = _to_dev(torch, src)
    d_dst = torch.empty_like(d_src)
    L.check(lib.cqb_ntt_bn254_fr_batch_map_dev(ctypes.c_void_p(d_src.data_ptr()), ctypes.c_void_p(d_dst.data_ptr()), L.p64(w), log_n, batch, seg_log, 1,
                                               L.p64(wb), big_log, row0, 0))
    L.check(lib.cqb_sync())
    assert np.array_equal(_to_host(d_dst).reshape(n, batch, 4), exp_t)

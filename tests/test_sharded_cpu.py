"""world_size-2 (and 3) gloo test of the multi-GPU host logic on CPU: point-range sharding, all-gather of the affine
partials, fold. The device operations are injected (an oracle-backed stand-in), so this covers exactly the code
bench.py runs under torchrun, minus the CUDA kernels (those are covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import cqb200
    from oracle import oracle_lib as O
    from sha2_on_cq_halo2_b200.sharded import ShardedMSM, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class OracleBackend:  # stand-in device for the CPU test
        def __init__(self, bases):
            self.bases = bases

        def msm(self, scalars):
            _, aff = O.best_multiexp(scalars, self.bases, 1)
            return aff, int(not aff.any())

        def msm_sparse(self, idx, scalars):
            aff = O.sparse_commit(self.bases, idx, scalars) if len(idx) else np.zeros(8, np.uint64)
            return aff, int(not aff.any())

        def sum_affine(self, pts):
            acc = np.zeros(12, np.uint64)
            for p in pts:
                acc = O.g1_add_ja(acc, p)
            aff = O.g1_to_affine(acc)
            return aff, int(not aff.any())

    start, cnt = shard_range(n, rank, world)
    bases = O.synth_bases(0xC0FFEE, n, 2)
    scalars = O.synth_scalars(0x5EED0001, n)
    if n > 4:
        scalars[1] = 0
        bases[2] = 0
    sm = ShardedMSM(OracleBackend(bases[start:start + cnt]), rank, world)
    got = sm.msm(scalars[start:start + cnt])
    _, exp = O.best_multiexp(scalars, bases, 2)
    ok = bool(np.array_equal(got.to_affine(), exp))
    # a shard whose partial is the identity must fold correctly too
    z = ShardedMSM(OracleBackend(bases[start:start + cnt]), rank, world)
    zero = z.msm(np.zeros((cnt, 4), np.uint64))
    ok = ok and zero.is_identity
    # sparse CQ commitment over the same sharded set
    rng = np.random.default_rng(3)
    sidx = np.sort(rng.choice(n, n // 3, replace=False)).astype(np.uint32)
    ssc = O.synth_scalars(77, sidx.shape[0])
    got_s = ShardedMSM(OracleBackend(bases[start:start + cnt]), rank, world)
    got_s.backend.n = cnt
    sp = got_s.msm_sparse(sidx, ssc, start)
    ok = ok and bool(np.array_equal(sp.to_affine(), O.sparse_commit(bases, sidx, ssc)))
    q.put((rank, ok, start, cnt))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 301), (3, 64)])
def test_sharded_msm_gloo(world, n):
    import torch.multiprocessing as mp

    from oracle import oracle_lib as O

    O.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    covered = sorted((s, c) for _, _, s, c in res)
    assert covered[0][0] == 0 and sum(c for _, c in covered) == n
    for (s0, c0), (s1, _) in zip(covered, covered[1:]):
        assert s0 + c0 == s1


def test_shard_range_balanced():
    sys.path.insert(0, ROOT)
    import cqb200  # noqa: F401
    from sha2_on_cq_halo2_b200.sharded import shard_range

    for n in (0, 1, 7, 8, 1 << 24, (1 << 24) + 5):
        for world in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and sum(c for _, c in rs) == n
            assert max(c for _, c in rs) - min(c for _, c in rs) <= 1
            for (s0, c0), (s1, _) in zip(rs, rs[1:]):
                assert s0 + c0 == s1


def _ntt_worker(rank, world, port, log_n, q, fused=False):
    """ShardedNTT host logic (three distributed transposes, batched local transforms, twiddles) with an oracle-backed
    stand-in device: the result block of every rank must equal the oracle's best_fft of the whole vector"""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import cqb200  # noqa: F401
    from oracle import oracle_lib as O
    from oracle import pyref as P
    from sha2_on_cq_halo2_b200.sharded import ShardedNTT

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class OracleNttBackend:  # torch uint8 CPU tensors, 32 bytes per element
        def empty(self, nelem):
            return torch.empty(nelem * 32, dtype=torch.uint8)

        @staticmethod
        def _np(t):
            return t.numpy().view(np.uint64).reshape(-1, 4)

        def transpose(self, t, rows, cols):
            a = self._np(t).reshape(rows, cols, 4).transpose(1, 0, 2)
            return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))

        def ntt_batch(self, t, omega_limbs, log_n, batch):
            a = self._np(t).reshape(batch, 1 << log_n, 4)
            for b in range(batch):
                a[b] = O.best_fft(np.ascontiguousarray(a[b]), omega_limbs, log_n, 1)

        def mul_omega_powers(self, t, rows, cols, row0, omega_limbs, log_n):
            a = self._np(t).reshape(rows, cols, 4)
            w = P.fr_array_to_ints(np.asarray(omega_limbs)[None, :])[0]
            for r in range(rows):
                vals = P.fr_array_to_ints(a[r])
                a[r] = P.fr_array_from_ints([v * pow(w, (row0 + r) * c, P.R_MOD) % P.R_MOD for c, v in enumerate(vals)])

        def scale(self, t, nelem, factor_limbs):
            a = self._np(t)
            f = P.fr_array_to_ints(np.asarray(factor_limbs)[None, :])[0]
            a[:] = P.fr_array_from_ints([v * f % P.R_MOD for v in P.fr_array_to_ints(a)])

        def interleave(self, recv, world, q_local, p_local):
            return recv.view(world, q_local, p_local * 32).permute(1, 0, 2).contiguous().view(-1)

    class FusedOracleNttBackend(OracleNttBackend):
        """adds the mapped batched transform (cqb_ntt_bn254_fr_batch_map_dev's address maps, restated with numpy indexing), so
        that ShardedNTT takes the SAME fused path it takes on the GPUs: gather from the all-to-all buffer, twiddles, transposed
        store"""

        def ntt_batch_map(self, src, omega_limbs, log_n, batch, in_seg_log, tw_omega_limbs=None, tw_log_n=0, tw_row0=0, src_offset_elems=0,
                          in_batch_total=0):
            n, seg, total = 1 << log_n, 1 << in_seg_log, (in_batch_total or batch)
            a = self._np(src)
            idx = np.arange(n)
            out = np.zeros((n, batch, 4), np.uint64)
            for b in range(batch):
                off = src_offset_elems + (idx >> in_seg_log) * (total * seg) + b * seg + (idx & (seg - 1))
                res = O.best_fft(np.ascontiguousarray(a[off]), omega_limbs, log_n, 1)
                if tw_omega_limbs is not None:
                    w = P.fr_array_to_ints(np.asarray(tw_omega_limbs)[None, :])[0]
                    res = P.fr_array_from_ints([v * pow(w, (tw_row0 + b) * i, P.R_MOD) % P.R_MOD for i, v in enumerate(P.fr_array_to_ints(res))])
                out[:, b] = res
            return torch.from_numpy(out.view(np.uint8).reshape(-1))

    n = 1 << log_n
    per = n // world
    full = O.synth_scalars(0x5EED0002, n)
    sn = ShardedNTT(FusedOracleNttBackend() if fused else OracleNttBackend(), log_n, rank, world)
    mine = torch.from_numpy(np.ascontiguousarray(full[rank * per:(rank + 1) * per]).view(np.uint8).reshape(-1).copy())
    got = sn.forward(mine).numpy().view(np.uint64).reshape(-1, 4)
    exp = O.best_fft(full.copy(), P.int_to_limbs(P.to_mont(sn.omega, P.R_MOD)), log_n, 1)
    ok = np.array_equal(got, exp[rank * per:(rank + 1) * per])
    back = sn.inverse(torch.from_numpy(np.ascontiguousarray(got).view(np.uint8).reshape(-1).copy())).numpy().view(np.uint64).reshape(-1, 4)
    ok = ok and np.array_equal(back, full[rank * per:(rank + 1) * per])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n,fused", [(2, 5, False), (2, 8, False), (4, 7, False), (1, 6, False), (2, 7, True), (4, 8, True), (1, 5, True)])
def test_sharded_ntt_host_logic_gloo(world, log_n, fused):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ntt_worker, args=(r, world, port, log_n, q, fused)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res

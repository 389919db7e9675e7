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

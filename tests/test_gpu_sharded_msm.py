"""The point-range-sharded MSM and the sharded sparse CQ commitments with the REAL device
backend: world_size 2 and 3 as separate processes on the one GPU of the test box, every rank running libcqb200's kernels on its
shard (CudaBackend), the 64-byte partials all-gathered over gloo (NCCL refuses two ranks on one device) and folded by the
device kernel. Compared with the CPU oracle on the unsharded inputs. bench.py runs the same classes over NCCL at N > 1."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, table_n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import cqb200
    from oracle import oracle_lib as O
    from sha2_on_cq_halo2_b200.sharded import CudaBackend, ShardedMSM, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cqb200._lib.init(0)
    ok = {}
    # dense: both layouts
    bases = O.synth_bases(0xC0FFEE, n, 4)
    scalars = O.synth_scalars(0x5EED0001, n)
    scalars[1] = 0
    bases[2] = 0
    _, exp = O.best_multiexp(scalars, bases, 4)
    start, cnt = shard_range(n, rank, world)
    for pre in (False, True):
        sm = ShardedMSM(CudaBackend(bases_affine=bases[start:start + cnt], precompute=pre), rank, world)
        ok[f"dense_pre{int(pre)}"] = bool(np.array_equal(sm.msm(scalars[start:start + cnt]).to_affine(), exp))
    ok["zero"] = ShardedMSM(CudaBackend(bases_affine=bases[start:start + cnt]), rank, world).msm(np.zeros((cnt, 4), np.uint64)).is_identity
    # sparse CQ commitments (m, A, Q_A, A_0: static_lookup/prover.rs:167-170, 245-257) over a table SRS sharded by index range
    tb = O.synth_bases(0x7AB1E, table_n, 4)
    rng = np.random.default_rng(11)
    for frac, label in ((0.3, "sparse"), (1.0, "full_support"), (0.0, "empty")):
        m = int(table_n * frac)
        sidx = np.sort(rng.choice(table_n, m, replace=False)).astype(np.uint32)
        ssc = O.synth_scalars(77 + m, max(m, 1))[:m]
        t0, tc = shard_range(table_n, rank, world)
        for pre in (False, True):
            sp = ShardedMSM(CudaBackend(bases_affine=tb[t0:t0 + tc], precompute=pre), rank, world).msm_sparse(sidx, ssc, t0)
            want = O.sparse_commit(tb, sidx, ssc) if m else np.zeros(8, np.uint64)
            ok[f"{label}_pre{int(pre)}"] = bool(np.array_equal(sp.to_affine(), want))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,table_n", [(2, 70001, 1 << 16), (3, 5000, 4099)])
def test_sharded_dense_and_sparse_msm_on_device(world, n, table_n):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, table_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    assert len(res) == world
    for rank, ok in res:
        bad = [k for k, v in ok.items() if not v]
        assert not bad, (rank, bad)

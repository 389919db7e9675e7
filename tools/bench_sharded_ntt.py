"""tools/bench_sharded_ntt.py — the distributed four-step NTT (sharded.ShardedNTT) under torchrun on 2-8 GPUs of one box:
correctness of every rank's block against the single-GPU transform at a size all ranks can hold, then CUDA-event timing
(max over ranks) of the block-distributed forward transform at the given sizes. Rank 0 prints JSON lines."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import cqb200
from sha2_on_cq_halo2_b200.fields import fr_to_limbs
from sha2_on_cq_halo2_b200.sharded import CudaNttBackend, ShardedNTT

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = cqb200._lib
lib = L.init(local)
stream = torch.cuda.Stream(device=local)
torch.cuda.set_stream(stream)
L.check(lib.cqb_set_stream(ctypes.c_void_p(stream.cuda_stream)))
dev = f"cuda:{local}"
be = CudaNttBackend(dev)


def synth(start, count):
    t = torch.empty(count * 32, dtype=torch.uint8, device=dev)
    L.check(lib.cqb_synth_scalars_dev(0x5EED0002, start, count, ctypes.c_void_p(t.data_ptr())))
    return t


modes = ["nccl"] + (["overlap", "peer"] if world > 1 else [])
if len(sys.argv) > 2:
    modes = [m for m in modes if m in sys.argv[2].split(",")]
# ---- correctness: every rank's block equals the single-GPU transform of the whole vector -------------------------------
for log_n, mode in [(l, m) for l in (10, 16, 20) for m in modes]:
    n = 1 << log_n
    per = n // world
    sn = ShardedNTT(be, log_n, rank, world)
    if mode == "peer":
        sn.enable_peer_exchange()
    if mode == "overlap":
        sn.enable_overlap(int(os.environ.get("NTT_GROUPS", "2")))
    got = sn.forward(synth(rank * per, per))
    full = synth(0, n)
    L.check(lib.cqb_ntt_bn254_fr_dev(ctypes.c_void_p(full.data_ptr()), L.p64(fr_to_limbs(sn.omega)), log_n))
    torch.cuda.synchronize()
    assert torch.equal(got, full[rank * per * 32:(rank + 1) * per * 32]), f"rank {rank}: block mismatch at 2^{log_n}"
    back = sn.inverse(got)
    torch.cuda.synchronize()
    assert torch.equal(back, synth(rank * per, per)), f"rank {rank}: inverse mismatch at 2^{log_n}"
    del full, got, back
    if mode == "peer":
        dist.barrier()
        be.release_peer_buffers()
if rank == 0:
    print(json.dumps({"check": "blocks of the distributed transform == single-GPU transform, inverse round trip", "sizes": [10, 16, 20],
                      "n_gpus": world, "ok": True}), flush=True)

# ---- timing ---------------------------------------------------------------------------------------------------------------
for log_n, mode in [(int(x), m) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["24", "26"]) for m in modes]:
    n = 1 << log_n
    per = n // world
    sn = ShardedNTT(be, log_n, rank, world)
    if mode == "peer":
        sn.enable_peer_exchange()
    if mode == "overlap":
        sn.enable_overlap(int(os.environ.get("NTT_GROUPS", "2")))
    x = synth(rank * per, per)
    for _ in range(3):
        y = sn.forward(x)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        y = sn.forward(x)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"log_n": log_n, "n_gpus": world, "distributed_forward_ms": round(float(ms.item()), 3),
                          "gelem_per_s": round(n / float(ms.item()) / 1e6, 3), "exchange": mode,
                          "layout": "block-distributed natural order in and out"}), flush=True)
    del x, y
    if mode == "peer":
        dist.barrier()
        be.release_peer_buffers()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

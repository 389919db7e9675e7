"""tools/a2a_probe.py — measurement behind the "NTT stays per-GPU" decision: NCCL all-to-all time for the
exchange a distributed four-step NTT of 2^log_n Fr elements needs (every rank sends (N-1)/N of its n/N x 32 B), next to the
local NTT time of one rank's share. Run under torchrun; rank 0 prints one JSON line per size."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import cqb200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = cqb200._lib
lib = L.init(local)
stream = torch.cuda.Stream(device=local)  # a real (non-NULL) stream shared by torch's events / NCCL and the library
torch.cuda.set_stream(stream)
L.check(lib.cqb_set_stream(ctypes.c_void_p(stream.cuda_stream)))
from sha2_on_cq_halo2_b200.fields import FR_ROOT_OF_UNITY, FR_S, R_MOD, fr_to_limbs

for log_n in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["24", "26"])]:
    n = 1 << log_n
    per = n // world
    src = torch.empty(per * 32, dtype=torch.uint8, device="cuda")
    dst = torch.empty_like(src)
    L.check(lib.cqb_synth_scalars_dev(1 + rank, 0, per, ctypes.c_void_p(src.data_ptr())))

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    a2a_ms = timed(lambda: dist.all_to_all_single(dst, src))
    # local transform of one rank's share (2^(log_n - log2 world) elements: the per-rank butterfly work is 1/N of the whole up to
    # the log factor) and the single-GPU transform of the whole vector for comparison (rank 0 only allocates it)
    lk = log_n - (world.bit_length() - 1)
    w = FR_ROOT_OF_UNITY
    for _ in range(lk, FR_S):
        w = w * w % R_MOD
    om = fr_to_limbs(w)
    local_ms = timed(lambda: L.check(lib.cqb_ntt_bn254_fr_dev(ctypes.c_void_p(src.data_ptr()), L.p64(om), lk)))
    if rank == 0:
        print(json.dumps({"log_n": log_n, "n_gpus": world, "all_to_all_ms": round(a2a_ms, 3), "bytes_per_rank": per * 32,
                          "local_ntt_2^%d_ms" % lk: round(local_ms, 3),
                          "projected_distributed_ms_3_exchanges": round(3 * a2a_ms + local_ms * log_n / lk, 3)}), flush=True)
    del src, dst
dist.barrier()
dist.destroy_process_group()

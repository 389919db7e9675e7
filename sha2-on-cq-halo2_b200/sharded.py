"""Point-range-sharded MSM across the GPUs of one box.

The reference splits a large multiexp into contiguous point ranges across rayon threads and folds the partial results
with Jacobian additions (halo2_proofs/src/arithmetic.rs:137-153). The multi-GPU form is the same decomposition with one
process per GPU: every rank keeps its range of the SRS resident in HBM, receives only its range of the scalars, computes
one partial point, and the `world` affine partials (64 B + flag each) are all-gathered (NCCL over NVLink on the GPU box,
gloo in the CPU tests) and folded by every rank. No other data-path collective exists on this path.

`backend` abstracts the two device operations so that the host logic can be exercised without a GPU (the CPU tests inject
an oracle-backed backend as a stand-in device; the product backend is CudaBackend = libcqb200.so, no fallback).
"""
import ctypes

import numpy as np

from . import _lib
from .arithmetic import G1


def shard_range(n, rank, world):
    """contiguous, balanced point ranges: the first n % world ranks own one extra point"""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


class CudaBackend:
    """device operations through the C ABI"""

    def __init__(self, bases_affine=None, device_ptr=None, n=None, precompute=False, window_bits=0):
        lib = _lib.lib()
        h = ctypes.c_uint64(0)
        if device_ptr is not None:
            _lib.check(lib.cqb_bases_register_device(ctypes.c_void_p(device_ptr), n, ctypes.byref(h)))
            self.n = n
        else:
            bases_affine = np.ascontiguousarray(bases_affine, dtype=np.uint64)
            self.n = bases_affine.shape[0]
            _lib.check(lib.cqb_bases_register(_lib.p64(bases_affine), self.n, ctypes.byref(h)))
        self.handle = h.value
        if precompute:
            _lib.check(lib.cqb_bases_precompute(self.handle, window_bits))
        self._out = np.zeros(8, np.uint64)
        self._inf = ctypes.c_int(0)

    def msm(self, scalars):
        """scalars: (m,4) uint64 host array, or an int device pointer with .count given via msm_dev"""
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        _lib.check(_lib.lib().cqb_msm_bn254_g1(self.handle, 0, _lib.p64(scalars), scalars.shape[0], _lib.p64(self._out),
                                               ctypes.byref(self._inf)))
        return self._out.copy(), self._inf.value

    def msm_host_ptr(self, host_ptr, m):
        _lib.check(_lib.lib().cqb_msm_bn254_g1(self.handle, 0, ctypes.cast(ctypes.c_void_p(host_ptr), _lib.u64p), m,
                                               _lib.p64(self._out), ctypes.byref(self._inf)))
        return self._out.copy(), self._inf.value

    def msm_dev(self, device_ptr, m):
        _lib.check(_lib.lib().cqb_msm_bn254_g1_dev(self.handle, 0, ctypes.c_void_p(device_ptr), m, _lib.p64(self._out),
                                                   ctypes.byref(self._inf)))
        return self._out.copy(), self._inf.value

    def msm_sparse(self, idx, scalars):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        _lib.check(_lib.lib().cqb_msm_bn254_g1_sparse(self.handle, idx.ctypes.data_as(_lib.u32p), _lib.p64(scalars), idx.shape[0],
                                                      _lib.p64(self._out), ctypes.byref(self._inf)))
        return self._out.copy(), self._inf.value

    def sum_affine(self, points):
        points = np.ascontiguousarray(points, dtype=np.uint64)
        out = np.zeros(8, np.uint64)
        inf = ctypes.c_int(0)
        _lib.check(_lib.lib().cqb_g1_sum_affine(_lib.p64(points), points.shape[0], _lib.p64(out), ctypes.byref(inf)))
        return out, inf.value


class ShardedMSM:
    """One rank's view of an MSM sharded by point range over `world` ranks."""

    def __init__(self, backend, rank=0, world=1, group=None, device="cpu"):
        self.backend = backend
        self.rank, self.world, self.group, self.device = rank, world, group, device
        if world > 1:
            import torch

            self._torch = torch
            self._in = torch.zeros(8, dtype=torch.int64, device=device)
            self._out = torch.zeros(8 * world, dtype=torch.int64, device=device)

    def fold(self, partial_affine):
        """all-gather the per-rank affine partials and add them up: arithmetic.rs:153 across GPUs. The identity is the
        all-zero point (derive/curve.rs:696-709), so no separate flag has to travel."""
        if self.world == 1:
            return G1(partial_affine, not partial_affine.any())
        import torch.distributed as dist

        torch = self._torch
        self._in.copy_(torch.from_numpy(np.ascontiguousarray(partial_affine).view(np.int64)))
        dist.all_gather_into_tensor(self._out, self._in, group=self.group)
        parts = self._out.cpu().numpy().view(np.uint64).reshape(self.world, 8)
        out, inf = self.backend.sum_affine(np.ascontiguousarray(parts))
        return G1(out, inf)

    def msm(self, local_scalars):
        partial, _ = self.backend.msm(local_scalars)
        return self.fold(partial)

    def msm_sparse(self, idx, scalars, shard_start):
        """CQ sparse commitments (m, A, Q_A, A_0) over a table SRS sharded by index range: this rank keeps
        the entries whose table index falls in its shard [shard_start, shard_start + backend.n) and rebases them"""
        idx = np.asarray(idx, dtype=np.int64)
        keep = (idx >= shard_start) & (idx < shard_start + self.backend.n)
        partial, _ = self.backend.msm_sparse((idx[keep] - shard_start).astype(np.uint32), np.ascontiguousarray(scalars)[keep])
        return self.fold(partial)

    def msm_dev(self, device_ptr, m):
        partial, _ = self.backend.msm_dev(device_ptr, m)
        return self.fold(partial)

    def msm_host_ptr(self, host_ptr, m):
        partial, _ = self.backend.msm_host_ptr(host_ptr, m)
        return self.fold(partial)

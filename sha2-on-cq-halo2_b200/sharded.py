"""Multi-GPU forms of the two hot operations on one box: the point-range-sharded MSM (below) and the
distributed four-step NTT (ShardedNTT, second half of this file).

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

    def msm_dev_to(self, device_ptr, m, out_device_ptr):
        """result left on the device (80 B), not waited for"""
        _lib.check(_lib.lib().cqb_msm_bn254_g1_dev_to(self.handle, 0, ctypes.c_void_p(device_ptr), m, ctypes.c_void_p(out_device_ptr)))

    def msm_host_ptr_to(self, host_ptr, m, out_device_ptr):
        _lib.check(_lib.lib().cqb_msm_bn254_g1_to(self.handle, 0, ctypes.cast(ctypes.c_void_p(host_ptr), _lib.u64p), m,
                                                  ctypes.c_void_p(out_device_ptr)))

    def msm_sparse(self, idx, scalars):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
        _lib.check(_lib.lib().cqb_msm_bn254_g1_sparse(self.handle, idx.ctypes.data_as(_lib.u32p), _lib.p64(scalars), idx.shape[0],
                                                      _lib.p64(self._out), ctypes.byref(self._inf)))
        return self._out.copy(), self._inf.value

    def sum_affine_dev(self, device_ptr, count):
        """fold `count` affine points that already sit in device memory (the all-gather's output): no host bounce"""
        out = np.zeros(8, np.uint64)
        inf = ctypes.c_int(0)
        _lib.check(_lib.lib().cqb_g1_sum_affine_dev(ctypes.c_void_p(device_ptr), count, _lib.p64(out), ctypes.byref(inf)))
        return out, inf.value

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
            self._raw = torch.zeros(10, dtype=torch.int64, device=device)  # 80-byte device result of the *_to calls

    def fold(self, partial_affine):
        """all-gather the per-rank affine partials and add them up: arithmetic.rs:153 across GPUs. The identity is the
        all-zero point (derive/curve.rs:696-709), so no separate flag has to travel."""
        if self.world == 1:
            return G1(partial_affine, not partial_affine.any())
        import torch.distributed as dist

        torch = self._torch
        self._in.copy_(torch.from_numpy(np.ascontiguousarray(partial_affine).view(np.int64)))
        dist.all_gather_into_tensor(self._out, self._in, group=self.group)
        if self._out.is_cuda and hasattr(self.backend, "sum_affine_dev"):
            # the library's kernels run on torch's current stream (cqb_set_stream), so the fold is ordered after the all-gather
            out, inf = self.backend.sum_affine_dev(self._out.data_ptr(), self.world)
            return G1(out, inf)
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

    def _fold_device(self):
        """the rank's partial is in self._raw on the device: all-gather its 64 point bytes and fold, one read-back in all"""
        import torch.distributed as dist

        dist.all_gather_into_tensor(self._out, self._raw[:8], group=self.group)
        out, inf = self.backend.sum_affine_dev(self._out.data_ptr(), self.world)
        return G1(out, inf)

    def _device_path(self):
        return self.world > 1 and self._out.is_cuda and hasattr(self.backend, "msm_dev_to")

    def msm_dev(self, device_ptr, m):
        if self._device_path():
            self.backend.msm_dev_to(device_ptr, m, self._raw.data_ptr())
            return self._fold_device()
        partial, _ = self.backend.msm_dev(device_ptr, m)
        return self.fold(partial)

    def msm_host_ptr(self, host_ptr, m):
        if self._device_path():
            self.backend.msm_host_ptr_to(host_ptr, m, self._raw.data_ptr())
            return self._fold_device()
        partial, _ = self.backend.msm_host_ptr(host_ptr, m)
        return self.fold(partial)


# ---------------------------------------------------------------------------------------------------------------------
# Distributed four-step NTT (SURVEY.md §8e "NTT: natural only via four-step"): measured to win on one NVSwitch box
# (tools/a2a_probe.py: the all-to-all of a 2^26-element vector over 8 GPUs takes 0.45 ms, the single-GPU transform 16.7 ms).
#
# The vector of n = n1 * n2 elements is block-distributed: rank g holds the contiguous range [g n/G, (g+1) n/G) on input AND
# on output (natural order both ways — the layout a caller that shards polynomials by index range already has). With
# j = j1 n2 + j2 and k = k1 + n1 k2:   X[k1 + n1 k2] = sum_j2 w_n2^(j2 k2) [ w^(j2 k1) sum_j1 x[j1 n2 + j2] w_n1^(j1 k1) ]
#   T1  distributed transpose  [j1][j2] -> [j2][j1]           (local 32-byte-element transpose + all-to-all + interleave)
#   A   n2/G local transforms of length n1 over j1 (batched)  -> [j2][k1]
#   B   multiply by w^(j2 k1)                                  (cqb_fr_mul_omega_powers_dev)
#   T2  distributed transpose  -> [k1][j2]
#   C   n1/G local transforms of length n2 over j2             -> [k1][k2]
#   T3  distributed transpose  -> [k2][k1] = natural order k = k2 n1 + k1, block-distributed
# The only data-path collective is the all-to-all of the three transposes (NCCL over NVLink; gloo in the CPU test).
# Field arithmetic is exact, so the result has the limbs of the reference's best_fft on the whole vector.
# ---------------------------------------------------------------------------------------------------------------------
class CudaNttBackend:
    """device operations of ShardedNTT through the C ABI, on torch uint8 CUDA tensors of 32 bytes per element"""

    def __init__(self, device):
        import torch

        self.torch = torch
        self.device = device
        self.bind_stream()

    def bind_stream(self):
        """The library's kernels and torch's (NCCL collectives, permute / copy kernels, the caching allocator) must be ordered
        on ONE stream: hand torch's current stream to the library (cqb_set_stream). Called at construction and again at the
        top of every distributed transform, so a caller that never touched cqb_set_stream — or switched torch streams in
        between — cannot race the all-to-all against the transpose / batched NTT kernels."""
        st = self.torch.cuda.current_stream(self.device).cuda_stream or 1  # 0 = torch's default stream = cudaStreamLegacy (handle 0x1)
        if getattr(self, "_bound", None) != st:
            _lib.check(_lib.lib().cqb_set_stream(ctypes.c_void_p(st)))
            self._bound = st

    def empty(self, nelem):
        return self.torch.empty(nelem * 32, dtype=self.torch.uint8, device=self.device)

    def transpose(self, t, rows, cols):
        out = self.empty(rows * cols)
        _lib.check(_lib.lib().cqb_fr_transpose_dev(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(out.data_ptr()), rows, cols))
        return out

    def ntt_batch(self, t, omega_limbs, log_n, batch):
        done = 0
        while done < batch:  # gridDim.y limit of the batched kernel
            b = min(65535, batch - done)
            _lib.check(_lib.lib().cqb_ntt_bn254_fr_batch_dev(ctypes.c_void_p(t.data_ptr() + (done << log_n) * 32), _lib.p64(omega_limbs), log_n, b))
            done += b

    def ntt_batch_map(self, src, omega_limbs, log_n, batch, in_seg_log, tw_omega_limbs=None, tw_log_n=0, tw_row0=0, src_offset_elems=0,
                      in_batch_total=0):
        """batched transform out of `src` (an all-to-all receive buffer [source rank][member][segment]) into a new buffer laid
        out [idx][member] — the gather, the twiddle step and the transposition are fused into the kernel's first and last pass"""
        assert batch <= 65535
        out = self.empty(batch << log_n)
        _lib.check(_lib.lib().cqb_ntt_bn254_fr_batch_map_dev(
            ctypes.c_void_p(src.data_ptr() + src_offset_elems * 32), ctypes.c_void_p(out.data_ptr()), _lib.p64(omega_limbs), log_n, batch, in_seg_log, 1,
            _lib.p64(tw_omega_limbs) if tw_omega_limbs is not None else None, tw_log_n, tw_row0, in_batch_total))
        return out

    # ---- peer-memory exchange (CUDA IPC): receive buffers every rank exposes to the others ---------------------------------
    class _RawBuffer:
        """a cqb_dev_alloc'ed buffer viewed as a torch tensor through the CUDA array interface"""

        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    def peer_buffers(self, nelem, count, rank, world, group):
        """allocate `count` receive buffers of nelem elements, exchange their IPC handles; returns (own tensors, per buffer the
        list of `world` device pointers: peers' buffers opened through CUDA IPC, own pointer at index rank)"""
        import torch.distributed as dist

        lib = _lib.lib()
        own_ptrs, own_tensors, handles = [], [], []
        for _ in range(count):
            d = ctypes.c_void_p()
            _lib.check(lib.cqb_dev_alloc(nelem * 32, ctypes.byref(d)))
            h = ctypes.create_string_buffer(64)
            _lib.check(lib.cqb_ipc_export(d, h))
            own_ptrs.append(d.value)
            handles.append(h.raw)
            own_tensors.append(self.torch.as_tensor(self._RawBuffer(d.value, nelem * 32), device=self.device))
        gathered = [None] * world
        dist.all_gather_object(gathered, handles, group=group)
        tables, opened = [], []
        for c in range(count):
            ptrs = []
            for r in range(world):
                if r == rank:
                    ptrs.append(own_ptrs[c])
                else:
                    p = ctypes.c_void_p()
                    _lib.check(lib.cqb_ipc_open(gathered[r][c], ctypes.byref(p)))
                    ptrs.append(p.value)
                    opened.append(p.value)
            tables.append(ptrs)
        self._ipc_state = getattr(self, "_ipc_state", []) + [(own_ptrs, opened)]
        return own_tensors, tables

    def release_peer_buffers(self):
        lib = _lib.lib()
        for own_ptrs, opened in getattr(self, "_ipc_state", []):
            for p in opened:
                _lib.check(lib.cqb_ipc_close(ctypes.c_void_p(p)))
            for p in own_ptrs:
                _lib.check(lib.cqb_dev_free(ctypes.c_void_p(p)))
        self._ipc_state = []

    def ntt_batch_p2p(self, src, peer_ptrs, rank, omega_limbs, log_n, batch, in_seg_log, tw_omega_limbs=None, tw_log_n=0, tw_row0=0):
        """the batched transform whose last pass stores into the peers' receive buffers (the exchange is the store)"""
        assert batch <= 65535
        scratch = self.empty(32)  # the library keeps its own scratch for the intermediate passes; d_scratch only has to differ from src
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[ctypes.c_void_p(p) for p in peer_ptrs])
        _lib.check(_lib.lib().cqb_ntt_bn254_fr_batch_p2p_dev(
            ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(scratch.data_ptr()), arr, len(peer_ptrs), rank, _lib.p64(omega_limbs), log_n, batch,
            in_seg_log, _lib.p64(tw_omega_limbs) if tw_omega_limbs is not None else None, tw_log_n, tw_row0))

    def mul_omega_powers(self, t, rows, cols, row0, omega_limbs, log_n):
        _lib.check(_lib.lib().cqb_fr_mul_omega_powers_dev(ctypes.c_void_p(t.data_ptr()), rows, cols, row0, _lib.p64(omega_limbs), log_n))

    def scale(self, t, nelem, factor_limbs):
        _lib.check(_lib.lib().cqb_fr_scale_dev(ctypes.c_void_p(t.data_ptr()), nelem, _lib.p64(factor_limbs)))

    def interleave(self, recv, world, q_local, p_local):
        """recv = [src rank][q_local][p_local] -> [q_local][src rank][p_local] (rows of the transposed matrix, whole)"""
        return recv.view(world, q_local, p_local * 32).permute(1, 0, 2).contiguous().view(-1)


class ShardedNTT:
    """One rank's view of a forward / inverse NTT of 2^log_n elements block-distributed over `world` ranks."""

    def __init__(self, backend, log_n, rank=0, world=1, group=None):
        from .fields import FR_ROOT_OF_UNITY, FR_S, R_MOD, fr_to_limbs

        assert world & (world - 1) == 0, "world must be a power of two"
        self.backend, self.rank, self.world, self.group = backend, rank, world, group
        self.log_n = log_n
        self.l1 = log_n // 2            # n1 = 2^l1 rows (j1 / k1), n2 = 2^l2 columns (j2 / k2)
        self.l2 = log_n - self.l1
        g = world.bit_length() - 1
        assert self.l1 >= g and self.l2 >= g, "every rank needs at least one row and one column"
        assert log_n <= FR_S
        w = FR_ROOT_OF_UNITY
        for _ in range(log_n, FR_S):
            w = w * w % R_MOD
        self._mod = R_MOD
        self._limbs = fr_to_limbs
        self.omega = w                                   # the 2^log_n-th root the reference's domain would use (domain.rs:54-61)

    def _dist_transpose(self, x, p_rows, q_cols):
        """global (p_rows x q_cols) matrix distributed by row blocks -> its transpose, distributed by row blocks"""
        import torch.distributed as dist

        G = self.world
        pl, ql = p_rows // G, q_cols // G
        t = self.backend.transpose(x, pl, q_cols)        # [q][p_local]; rows q in block h are contiguous: chunk h
        if G == 1:
            return t
        recv = self.backend.empty(q_cols * pl)
        dist.all_to_all_single(recv, t, group=self.group)
        return self.backend.interleave(recv, G, ql, pl)  # [q_local][p]

    def _all_to_all(self, t):
        if self.world == 1:
            return t
        import torch.distributed as dist

        recv = self.backend.empty(t.numel() // 32)
        dist.all_to_all_single(recv, t, group=self.group)
        return recv

    def _run_fused(self, x_local, omega):
        """the same six steps with every layout change but the first transposition and the last interleave folded into the
        batched transforms (cqb_ntt_bn254_fr_batch_map_dev): 2 memory passes besides the transforms instead of 7"""
        n1, n2, G = 1 << self.l1, 1 << self.l2, self.world
        pl, ql = n1 // G, n2 // G
        lim, be = self._limbs, self.backend
        t = be.transpose(x_local, pl, n2)                                            # [j2][j1_local]: chunk h = rows j2 of rank h
        r1 = self._all_to_all(t)                                                     # [src][j2_local][j1_local]
        a = be.ntt_batch_map(r1, lim(pow(omega, n2, self._mod)), self.l1, ql, pl.bit_length() - 1,
                             lim(omega), self.log_n, self.rank * ql)                 # A + B -> [k1][j2_local]: chunk h = rows k1 of rank h
        r2 = self._all_to_all(a)                                                     # [src][k1_local][j2_local(src)]
        c = be.ntt_batch_map(r2, lim(pow(omega, n1, self._mod)), self.l2, pl, ql.bit_length() - 1)  # C -> [k2][k1_local]
        r3 = self._all_to_all(c)                                                     # [src][k2_local][k1_local(src)]
        return be.interleave(r3, G, ql, pl) if G > 1 else r3                         # [k2_local][k1]: natural order

    def enable_peer_exchange(self):
        """switch to the peer-memory exchange: two receive buffers per rank, opened by every peer through CUDA IPC. After this,
        the second and third exchange are the last store of the batched transforms (cqb_ntt_bn254_fr_batch_p2p_dev) and only a
        stream-ordered one-element all-reduce separates the steps."""
        assert self.world > 1 and hasattr(self.backend, "peer_buffers")
        per = (1 << self.log_n) // self.world
        self._recv, self._peer_tables = self.backend.peer_buffers(per, 2, self.rank, self.world, self.group)
        self._flag = self.backend.torch.zeros(1, dtype=self.backend.torch.int32, device=self.backend.device)
        self._p2p = True

    def _rank_sync(self):
        import torch.distributed as dist

        dist.all_reduce(self._flag, group=self.group)  # stream-ordered: completes when every rank's preceding kernel has finished

    def _run_p2p(self, x_local, omega):
        n1, n2, G = 1 << self.l1, 1 << self.l2, self.world
        pl, ql = n1 // G, n2 // G
        lim, be = self._limbs, self.backend
        t = be.transpose(x_local, pl, n2)
        r1 = self._all_to_all(t)
        self._rank_sync()                       # nobody is still reading the receive buffers of the previous call
        be.ntt_batch_p2p(r1, self._peer_tables[0], self.rank, lim(pow(omega, n2, self._mod)), self.l1, ql, pl.bit_length() - 1,
                         lim(omega), self.log_n, self.rank * ql)
        self._rank_sync()                       # every rank's stores into my buffer 0 have landed
        be.ntt_batch_p2p(self._recv[0], self._peer_tables[1], self.rank, lim(pow(omega, n1, self._mod)), self.l2, pl, ql.bit_length() - 1)
        self._rank_sync()
        return be.interleave(self._recv[1], G, ql, pl)

    def enable_overlap(self, groups=2):
        """split the batched transforms into `groups` member groups: the all-to-all of a finished group (NCCL on a second
        stream, list form, chunks placed [source rank][group]) runs under the transform of the next group"""
        assert self.world > 1 and hasattr(self.backend, "ntt_batch_map")
        assert groups >= 1 and (groups & (groups - 1)) == 0, "member groups must be a power of two (segment logs are taken from them)"
        torch = self.backend.torch
        self._groups = groups
        self._comm_stream = torch.cuda.Stream(device=self.backend.device)

    def _run_overlap(self, x_local, omega):
        import torch.distributed as dist

        torch = self.backend.torch
        n1, n2, G, S = 1 << self.l1, 1 << self.l2, self.world, self._groups
        pl, ql = n1 // G, n2 // G
        lim, be = self._limbs, self.backend
        main = torch.cuda.current_stream()
        comm = self._comm_stream
        t = be.transpose(x_local, pl, n2)
        r1 = self._all_to_all(t)                                   # [src][j2_local][j1_local]

        def stage(src, length_log, members, seg, tw):
            """transform `members` members in S groups out of src = [G * S' segments][members][seg]; returns the receive buffer
            laid out [src rank][group][rows of this rank][members / S]"""
            gm = members // S
            rows = (1 << length_log) // G
            recv = be.empty(rows * members * G)
            rv = recv.view(G, S, rows * gm * 32)
            outs, evs = [], []
            for s in range(S):
                o = be.ntt_batch_map(src, lim(pow(omega, (1 << self.log_n) >> length_log, self._mod)), length_log, gm, seg.bit_length() - 1,
                                     lim(omega) if tw else None, self.log_n if tw else 0, self.rank * members + s * gm if tw else 0,
                                     src_offset_elems=s * gm * seg, in_batch_total=members)
                outs.append(o)                                   # [idx][gm]: chunk h = rows idx of rank h
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(comm):
                    comm.wait_event(ev)
                    dist.all_to_all([rv[g][s] for g in range(G)], list(o.view(G, rows * gm * 32).unbind(0)), group=self.group)
                    done = torch.cuda.Event()
                    done.record(comm)
                evs.append(done)
            for e in evs:
                main.wait_event(e)
            return recv, outs

        r2, keep_a = stage(r1, self.l1, ql, pl, True)             # A + B; r2 = [src][s][k1_local][ql / S]
        r3, keep_c = stage(r2, self.l2, pl, ql // S, False)       # C;     r3 = [src][s][k2_local][pl / S]
        out = r3.view(G, S, ql, (pl // S) * 32).permute(2, 0, 1, 3).contiguous().view(-1)   # [k2_local][k1]
        del keep_a, keep_c
        return out

    def _run(self, x_local, omega):
        n1, n2, G = 1 << self.l1, 1 << self.l2, self.world
        if hasattr(self.backend, "bind_stream"):
            self.backend.bind_stream()
        S = getattr(self, "_groups", 0)
        if S > 1 and min(n1, n2) // G >= S and (n1 // G) % S == 0 and (n2 // G) % S == 0 and max(n1, n2) // G <= 65535:
            return self._run_overlap(x_local, omega)
        if getattr(self, "_p2p", False) and max(n1, n2) // G <= 65535:
            return self._run_p2p(x_local, omega)
        if hasattr(self.backend, "ntt_batch_map") and max(n1, n2) // G <= 65535:
            return self._run_fused(x_local, omega)
        lim = self._limbs
        y = self._dist_transpose(x_local, n1, n2)                                    # T1: [j2_local][j1]
        self.backend.ntt_batch(y, lim(pow(omega, n2, self._mod)), self.l1, n2 // G)  # A
        self.backend.mul_omega_powers(y, n2 // G, n1, self.rank * (n2 // G), lim(omega), self.log_n)  # B
        z = self._dist_transpose(y, n2, n1)                                          # T2: [k1_local][j2]
        self.backend.ntt_batch(z, lim(pow(omega, n1, self._mod)), self.l2, n1 // G)  # C
        return self._dist_transpose(z, n1, n2)                                       # T3: [k2_local][k1] = natural order

    def forward(self, x_local):
        """best_fft(a, omega, log_n) (arithmetic.rs:171) on the distributed vector; returns this rank's block of the result"""
        return self._run(x_local, self.omega)

    def inverse(self, x_local):
        """EvaluationDomain::ifft (poly/domain.rs:366-374): best_fft with omega^-1, then the 1/n scaling"""
        out = self._run(x_local, pow(self.omega, -1, self._mod))
        self.backend.scale(out, (1 << self.log_n) // self.world, self._limbs(pow(1 << self.log_n, -1, self._mod)))
        return out

// msm.cu — BN254 (bn256) G1 multi-scalar multiplication, replacing best_multiexp / multiexp_serial
// (reference halo2_proofs/src/arithmetic.rs:13-159) behind every KZG commitment.
//
// The reference: unsigned c = ceil(ln n)-bit windows, one Jacobian bucket array per window filled by a serial loop,
// running-sum reduction, c doublings between windows, rayon chunks over point ranges. Only the affine normal form of the
// result is canonical, so any correct evaluation order yields identical output; this file evaluates the
// same sum the B200 way:
//   1. msm_count   : Montgomery -> canonical scalar (one modmul), SIGNED c-bit digits (2^(c-1) buckets per window, half
//                    the reference's bucket count), per-bucket histogram with L2 atomics;
//   2. msm_scan    : exclusive scan of the histogram -> bucket offsets;
//   3. msm_scatter : counting-sort scatter of (point id, sign) into bucket order;
//   4. msm_accumulate : the bucket-sorted lists are cut into equal chunks, one thread per chunk (load balance does not
//                    depend on the scalar distribution): XYZZ mixed additions (8M+2S) of its points, gathered from the
//                    device-resident SRS with 128-bit loads; negation folded into the load; all exceptional cases
//                    (identity base, P+P, P+(-P)) handled as the reference does (derive/curve.rs:866-871);
//                    msm_merge adds the per-chunk partials of buckets that straddle chunk boundaries;
//   5. msm_reduce  : sum_d d*B_d by chunked running sums (each thread: running-sum over its chunk, then a short
//                    double-and-add for the chunk offset), xyzz_sum: tree-add the chunk partials,
//   6. msm_final   : Horner over windows (c doublings each), normalise to affine, write x||y + identity flag.
//
// Two layouts share these kernels:
//   * WINDOWED (any bases, e.g. the one-shot host call): nwin = 254/c + 1 independent bucket sets, c <= 16.
//   * SINGLE SET over a PRECOMPUTED table (registered SRS; memory laid out for the 180 GB of a B200): the table holds
//     2^(c w) P_i for every window w (nwin x n x 64 B), so all windows feed ONE bucket set: larger c (up to 23) => fewer
//     windows => fewer bucket additions per point, one bucket reduction instead of nwin, no inter-window doublings.
// Roofline: integer pipe — 10 modmuls per bucket addition x nwin additions per point; the 64 B gather per addition is
// < 10 % of HBM bandwidth at that rate.
#include <algorithm>
#include <vector>

#include "internal.h"

namespace cqb {

static int g_forced_c = 0;
void msm_set_window_bits(int c) { g_forced_c = c; }

// -------------------------------------------------------------------------------------------------------------------
// device helpers
// -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fq ldg_fq(const uint4* p) {
    uint4 a = __ldg(p), b = __ldg(p + 1);
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fq(uint4* p, const Fq& v) {
    p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fq ld_fq(const uint4* p) {
    uint4 a = p[0], b = p[1];
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ G1Xyzz ld_xyzz(const uint4* p) {
    G1Xyzz r;
    r.x = ld_fq(p); r.y = ld_fq(p + 2); r.zz = ld_fq(p + 4); r.zzz = ld_fq(p + 6);
    return r;
}
__device__ __forceinline__ void st_xyzz(uint4* p, const G1Xyzz& v) {
    st_fq(p, v.x); st_fq(p + 2, v.y); st_fq(p + 4, v.zz); st_fq(p + 6, v.zzz);
}

// canonical 254-bit scalar -> raw c-bit digit of window w
__device__ __forceinline__ uint32_t window_bits(const uint32_t* k, int w, int c) {
    int bit = w * c;
    int limb = bit >> 5, sh = bit & 31;
    if (limb >= 8) return 0;
    uint64_t v = k[limb];
    if (limb + 1 < 8) v |= (uint64_t)k[limb + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

struct MsmShape {
    int c;            // window bits
    int nwin;         // number of digit windows per scalar
    int nsets;        // bucket sets: nwin (windowed) or the batch size B (one set per MSM over a precomputed table)
    int single;       // 1 = precomputed-table layout: all windows of an MSM share one bucket set
    uint32_t nb;      // buckets per set = 2^(c-1) (bucket ids 1..nb)
    uint32_t stride;  // nb + 2 : per-set stride of the histogram / offset arrays
    uint32_t table_n; // single set: points per window row of the precomputed table
    uint32_t offset;  // first base of this MSM inside the registered set
    size_t list_cap;  // capacity of one set's sorted list: n (windowed) or n * nwin (single set)
};

__device__ __forceinline__ Fr load_scalar_canonical(const uint4* scalars, size_t i) {
    uint4 a = __ldg(scalars + 2 * i), b = __ldg(scalars + 2 * i + 1);
    Fr k;
    k.l[0] = a.x; k.l[1] = a.y; k.l[2] = a.z; k.l[3] = a.w; k.l[4] = b.x; k.l[5] = b.y; k.l[6] = b.z; k.l[7] = b.w;
    return fp_from_mont<FrP>(k);  // reference arithmetic.rs:14 to_repr()
}

// Hot-bucket handling shared by the count and scatter kernels. Repeated scalars (all-equal columns, 0/1 witnesses, the
// identical upper digits of "negative small" values r - x) send a whole warp to ONE bucket; plain per-lane atomics would
// then serialise on one L2 address (16 M atomics on one counter at 2^24). A warp therefore first tests, with two ballots
// and a shuffle, whether the first non-zero digit is shared by at least DUP_MIN lanes; only then does it pay for
// __match_any_sync (MATCH runs on the ADU pipe at a fraction of the ballot rate and was the bound of the count kernel:
// 91 % ADU utilisation) and lets one leader per distinct bucket add the lanes' total. Uniform digits take the plain path.
constexpr int DUP_MIN = 4;
// 0: every lane has a zero digit (nothing to do) | 1: plain per-lane atomics | 2: aggregate with __match_any_sync.
// Lane 0's digit is the probe (one shuffle + one ballot per window). A zero probe shared by >= DUP_MIN lanes also takes
// the aggregated path: correct, just not the cheapest — zero digits are either rare (uniform scalars) or warp-wide (small
// scalars: every upper window is zero in all lanes and is skipped).
__device__ __forceinline__ int warp_key_mode(uint32_t key) {
    const uint32_t first = __shfl_sync(0xffffffffu, key, 0);
    const uint32_t same = __ballot_sync(0xffffffffu, key == first);
    if (first == 0u && same == 0xffffffffu) return 0;
    return __popc(same) >= DUP_MIN ? 2 : 1;
}

// (1) histogram of signed digits. One thread per scalar.
__global__ void __launch_bounds__(256) msm_count_kernel(const uint4* __restrict__ scalars, size_t n, MsmShape s,
                                                        uint32_t* __restrict__ hist) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t bset = blockIdx.y;  // batch member (single layout); 0 otherwise
    Fr k = active ? load_scalar_canonical(scalars + (size_t)bset * n * 2, i) : Fr::zero();
    uint32_t carry = 0;
    for (int w = 0; w < s.nwin; w++) {
        uint32_t d = window_bits(k.l, w, s.c) + carry;
        carry = 0;
        if (d > s.nb) { d = (1u << s.c) - d; carry = 1; }
        const uint32_t key = active ? d : 0u;
        const int mode = warp_key_mode(key);
        if (mode == 0) continue;  // zero digits contribute nothing (warp-uniform branch)
        uint32_t* h = hist + (s.single ? (size_t)bset : (size_t)w) * s.stride;
        if (mode == 2) {
            uint32_t peers = __match_any_sync(0xffffffffu, key);
            if (key && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(h + key, (uint32_t)__popc(peers));
        } else if (key) {
            atomicAdd(h + key, 1u);
        }
    }
}

// Besides the offsets, the scans leave in place of the histogram nxt[d] = the smallest NON-EMPTY bucket id >= d (NO_BUCKET
// if none): the accumulation kernel steps from one bucket to the next with a single load, however many empty ids lie
// between them (a handful of hot buckets 2^19 ids apart — repeated scalars, the constant upper digits of r - x — used to
// cost one thread a walk as long as the whole kernel).
constexpr uint32_t NO_BUCKET = 0xffffffffu;

// exclusive suffix minimum over the 1024 threads of a CTA: min of v over the threads with a HIGHER index
__device__ __forceinline__ uint32_t block_excl_suffix_min_1024(uint32_t v, uint32_t* warp_mins) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t y = __shfl_down_sync(0xffffffffu, x, off);
        if (lane + off < 32) x = min(x, y);
    }
    if (lane == 0) warp_mins[wid] = x;  // inclusive suffix min of the warp
    __syncthreads();
    if (wid == 0) {
        uint32_t m = warp_mins[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t y = __shfl_down_sync(0xffffffffu, m, off);
            if (lane + off < 32) m = min(m, y);
        }
        warp_mins[lane] = m;
    }
    __syncthreads();
    uint32_t e = __shfl_down_sync(0xffffffffu, x, 1);
    if (lane == 31) e = NO_BUCKET;
    uint32_t after = (wid < 31) ? warp_mins[wid + 1] : NO_BUCKET;
    __syncthreads();
    return min(e, after);
}

// (2a) exclusive scan, one CTA per bucket set (windowed layout: nb <= 32768)
__global__ void __launch_bounds__(1024) msm_scan_kernel(uint32_t* __restrict__ hist, uint32_t* __restrict__ offs,
                                                        uint32_t* __restrict__ cursor, MsmShape s, uint32_t* __restrict__ nonempty) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const int w = blockIdx.x;
    uint32_t* h = hist + (size_t)w * s.stride;
    uint32_t* o = offs + (size_t)w * s.stride;
    uint32_t* cu = cursor + (size_t)w * s.stride;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    uint32_t last_base = 1;
    for (uint32_t base = 1; base <= s.nb + 1; base += 1024) {
        last_base = base;
        uint32_t d = base + tid;
        uint32_t v = (d <= s.nb) ? h[d] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, ws, off);
                if (lane >= off) ws += y;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        uint32_t excl = carry_s + (wid ? warp_sums[wid - 1] : 0u) + (x - v);
        if (d <= s.nb + 1) { o[d] = excl; cu[d] = excl; }
        __syncthreads();
        if (tid == 1023) carry_s = excl + v;
        __syncthreads();
    }
    // backward pass: hist[d] <- nxt[d]; the number of non-empty buckets goes to *nonempty
    if (tid == 0) carry_s = NO_BUCKET;
    __syncthreads();
    uint32_t mine = 0;
    for (uint32_t base = last_base;; base -= 1024) {
        uint32_t d = base + tid;
        uint32_t v = (d <= s.nb && h[d] != 0u) ? d : NO_BUCKET;
        mine += (v != NO_BUCKET);
        uint32_t nx = min(v, min(block_excl_suffix_min_1024(v, warp_sums), carry_s));
        if (d <= s.nb) h[d] = nx;
        __syncthreads();
        if (tid == 0) carry_s = nx;
        __syncthreads();
        if (base == 1) break;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicAdd(nonempty, mine);
}

// (2b) three-kernel scan for one large bucket set (single-set layout, nb up to 2^22): tile sums, scan of the tile sums,
// per-tile scan with its offset. TILE = 1024 threads x 8 ids. tile_sums[0..ntiles) = sums -> offsets;
// tile_sums[ntiles..2 ntiles) = first non-empty id of the tile -> first non-empty id of any LATER tile.
constexpr uint32_t SCAN_TILE = 8192;
__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t* warp_sums, uint32_t* total) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = warp_sums[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, ws, off);
            if (lane >= off) ws += y;
        }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    uint32_t excl = (wid ? warp_sums[wid - 1] : 0u) + (x - v);
    if (total) *total = warp_sums[31];
    __syncthreads();
    return excl;
}
__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(const uint32_t* __restrict__ hist, uint32_t nb, uint32_t* __restrict__ tile_sums,
                                                              uint32_t ntiles) {
    __shared__ uint32_t ws[32];
    uint32_t base = 1 + blockIdx.x * SCAN_TILE + threadIdx.x * 8;
    uint32_t v = 0, first = NO_BUCKET;
#pragma unroll
    for (int j = 7; j >= 0; j--) {
        uint32_t d = base + j;
        if (d <= nb) { uint32_t c = hist[d]; v += c; if (c) first = d; }
    }
    uint32_t total;
    (void)block_excl_scan_1024(v, ws, &total);
    uint32_t later = block_excl_suffix_min_1024(first, ws);
    if (threadIdx.x == 0) {
        tile_sums[blockIdx.x] = total;
        tile_sums[ntiles + blockIdx.x] = min(first, later);
    }
}
__global__ void __launch_bounds__(1024) scan_tile_offsets_kernel(uint32_t* __restrict__ tile_sums, uint32_t ntiles) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    uint32_t last_base = 0;
    for (uint32_t base = 0; base < ntiles; base += 1024) {
        last_base = base;
        uint32_t t = base + threadIdx.x;
        uint32_t v = t < ntiles ? tile_sums[t] : 0u;
        uint32_t total;
        uint32_t excl = block_excl_scan_1024(v, ws, &total);
        if (t < ntiles) tile_sums[t] = carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    // first non-empty id of any later tile (exclusive suffix minimum), backward over the strips
    uint32_t* tf = tile_sums + ntiles;
    if (threadIdx.x == 0) carry = NO_BUCKET;
    __syncthreads();
    for (uint32_t base = last_base;; base -= 1024) {
        uint32_t t = base + threadIdx.x;
        uint32_t v = t < ntiles ? tf[t] : NO_BUCKET;
        uint32_t ex = min(block_excl_suffix_min_1024(v, ws), carry);
        if (t < ntiles) tf[t] = ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = min(v, ex);
        __syncthreads();
        if (base == 0) break;
    }
}
__global__ void __launch_bounds__(1024) scan_apply_kernel(uint32_t* __restrict__ hist, uint32_t nb, const uint32_t* __restrict__ tile_sums,
                                                          uint32_t ntiles, uint32_t* __restrict__ offs, uint32_t* __restrict__ cursor,
                                                          uint32_t* __restrict__ nonempty) {
    __shared__ uint32_t ws[32];
    uint32_t base = 1 + blockIdx.x * SCAN_TILE + threadIdx.x * 8;
    uint32_t vals[8], v = 0, first = NO_BUCKET;
#pragma unroll
    for (int j = 7; j >= 0; j--) {
        uint32_t d = base + j;
        vals[j] = (d <= nb) ? hist[d] : 0u;
        v += vals[j];
        if (vals[j]) first = d;
    }
    uint32_t excl = block_excl_scan_1024(v, ws, nullptr) + tile_sums[blockIdx.x];
    uint32_t run = min(block_excl_suffix_min_1024(first, ws), tile_sums[ntiles + blockIdx.x]);  // first non-empty id after this thread's 8
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t d = base + j;
        if (d <= nb + 1) { offs[d] = excl; cursor[d] = excl; }
        excl += vals[j];
    }
    uint32_t mine = 0;
#pragma unroll
    for (int j = 7; j >= 0; j--) {
        uint32_t d = base + j;
        if (vals[j]) { run = d; mine++; }
        if (d <= nb) hist[d] = run;  // nxt[d]
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(nonempty, mine);
}

// (3) scatter (point id, sign) into bucket order. One thread per scalar; digits are recomputed (1 modmul) instead of
// being stored and re-read (saves 2 x 4 B x nwin per point of HBM traffic). Windows are handled in groups of SCAT_G: the
// group's cursor atomics are all issued before the first of their results is consumed, so a warp waits for one L2
// round trip per group instead of one per window (the kernel was bound by exactly that wait: long-scoreboard stalls).
constexpr int SCAT_G = 4;
__global__ void __launch_bounds__(256) msm_scatter_kernel(const uint4* __restrict__ scalars, const uint32_t* __restrict__ idx,
                                                          size_t n, MsmShape s, uint32_t* __restrict__ cursor,
                                                          uint32_t* __restrict__ sorted, int range_shift) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t bset = blockIdx.y;
    // bucket-range passes (gridDim.z of them, executed one after the other: z is the slowest grid dimension): pass z only
    // places the digits whose bucket id lies in its range, so the number of write streams open at a time — one partially
    // written 128 B line per bucket — fits the L2 and the lines leave it complete
    const uint32_t my_range = blockIdx.z;
    Fr k = active ? load_scalar_canonical(scalars + (size_t)bset * n * 2, i) : Fr::zero();
    uint32_t pid = active ? (idx ? __ldg(idx + i) : (uint32_t)i + s.offset) : 0u;
    uint32_t carry = 0;
    for (int w0 = 0; w0 < s.nwin; w0 += SCAT_G) {
        uint32_t key[SCAT_G], neg[SCAT_G], base[SCAT_G], rank[SCAT_G], leader[SCAT_G];
#pragma unroll
        for (int g = 0; g < SCAT_G; g++) {
            const int w = w0 + g;
            key[g] = 0; neg[g] = 0; base[g] = 0; rank[g] = 0; leader[g] = 32;  // leader 32: this lane owns its own result
            if (w >= s.nwin) continue;
            uint32_t d = window_bits(k.l, w, s.c) + carry;
            carry = 0;
            if (d > s.nb) { d = (1u << s.c) - d; carry = 1; neg[g] = 1; }
            key[g] = (active && ((d - 1u) >> range_shift) == my_range) ? d : 0u;  // d == 0 wraps to a range that does not exist
            const int mode = warp_key_mode(key[g]);
            if (mode == 0) continue;
            uint32_t* cu = cursor + (s.single ? (size_t)bset : (size_t)w) * s.stride;
            if (mode == 2) {
                // one atomic per distinct bucket per warp, lanes take consecutive slots
                uint32_t peers = __match_any_sync(0xffffffffu, key[g]);
                leader[g] = (uint32_t)(__ffs(peers) - 1);
                rank[g] = (uint32_t)__popc(peers & ((1u << lane) - 1u));
                if (key[g] && lane == leader[g]) base[g] = atomicAdd(cu + key[g], (uint32_t)__popc(peers));
            } else if (key[g]) {
                base[g] = atomicAdd(cu + key[g], 1u);
            }
        }
#pragma unroll
        for (int g = 0; g < SCAT_G; g++) {
            const int w = w0 + g;
            if (w >= s.nwin) continue;
            // warp-uniform: leader[g] < 32 on every lane of a warp that took the aggregated path for this window
            if (__any_sync(0xffffffffu, leader[g] < 32u)) base[g] = __shfl_sync(0xffffffffu, base[g], leader[g] & 31u);
            if (key[g]) {
                uint32_t pos = base[g] + rank[g];
                if (s.single) sorted[(size_t)bset * s.list_cap + pos] = ((pid + (uint32_t)w * s.table_n) << 1) | neg[g];  // table row w
                else sorted[(size_t)w * s.list_cap + pos] = (pid << 1) | neg[g];
            }
        }
    }
}

// the largest d in [lo, hi] with o[d] <= pos
__device__ __forceinline__ uint32_t bucket_of_pos(const uint32_t* __restrict__ o, uint32_t lo, uint32_t hi, uint32_t pos) {
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (o[mid] <= pos) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// (4) bucket accumulation, load-balanced independently of the scalar distribution: each set's bucket-sorted list is
// cut into chunks of 2^seg_log entries and ONE THREAD OWNS ONE CHUNK (not one bucket), so every thread performs the
// same number of XYZZ mixed additions whether the digits are uniform, all equal, or (the top window) only a few bits
// wide. Within its chunk a thread walks the bucket boundaries (offs[]): buckets that start and end inside the chunk are
// complete and stored directly; the first and the last bucket of a chunk may continue in the neighbouring chunks and
// are stored as "head" / "tail" partials which msm_merge_kernel adds up (one XYZZ add per chunk boundary).
//
// Two instantiations are launched back to back and exactly one of them does the work, chosen on the device by how many
// buckets are occupied (no host round trip): DENSE steps to the next bucket id (almost every id is occupied — uniform
// scalars; this loop shape is 2.6 % faster there, 32.2 vs 33.0 ms at 2^24), SPARSE (< 1/4 of the ids occupied) follows the
// nxt table so that a step costs the same however many empty ids lie in between.
template <bool SPARSE>
__global__ void __launch_bounds__(128) msm_accumulate_kernel(const uint4* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                             const uint32_t* __restrict__ offs, const uint32_t* __restrict__ nxt,
                                                             const uint32_t* __restrict__ nonempty, MsmShape s, int seg_log, uint32_t cpw,
                                                             uint4* __restrict__ buckets, uint4* __restrict__ head, uint4* __restrict__ tail) {
    if (((size_t)*nonempty * 4 < (size_t)s.nb * s.nsets) != SPARSE) return;
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)s.nsets * cpw) return;
    uint32_t w = (uint32_t)(gid / cpw), k = (uint32_t)(gid % cpw);
    const uint32_t* o = offs + (size_t)w * s.stride;
    const uint32_t* nx = nxt + (size_t)w * s.stride;
    const uint32_t total = o[s.nb + 1];
    const uint32_t start = k << seg_log;
    if (start >= total) return;
    const uint32_t end = min(start + (1u << seg_log), total);
    // bucket d with o[d] <= start < o[d+1]: the largest d in [1, nb] with o[d] <= start
    uint32_t d = bucket_of_pos(o, 1, s.nb, start), bound = o[d + 1];
    bool is_first = true;
    const uint32_t* lst = sorted + (size_t)w * s.list_cap;
    G1Xyzz acc = G1Xyzz::identity();
    // ONE loop over the chunk's positions (all lanes of a warp stay in lockstep; a nested per-bucket loop diverges and was
    // measured 2x slower)
    for (uint32_t pos = start; pos < end; pos++) {
        if (pos >= bound) {  // bucket d ends inside this chunk
            if (is_first) { st_xyzz(head + gid * 8, acc); is_first = false; }
            else st_xyzz(buckets + ((size_t)w * s.nb + (d - 1)) * 8, acc);  // started and ended inside: complete
            acc = G1Xyzz::identity();
            if (!SPARSE) {
                do { d++; bound = o[d + 1]; } while (pos >= bound);
            } else {
                d = nx[d + 1];  // next non-empty bucket (it exists: pos < total)
                bound = o[d + 1];
            }
        }
        uint32_t e = __ldg(lst + pos);
        const uint4* bp = bases + (size_t)(e >> 1) * 4;
        Fq x = ldg_fq(bp), y = ldg_fq(bp + 2);
        if (x.is_zero() && y.is_zero()) continue;  // identity base contributes nothing (reference curve.rs:857-858)
        if (e & 1u) y = fp_neg<FqP>(y);
        g1_madd(acc, x, y);
    }
    if (is_first) st_xyzz(head + gid * 8, acc);
    else st_xyzz(tail + gid * 8, acc);
}

// (4') BATCHED-AFFINE bucket accumulation (the CPU form is the reference's batch_add, arithmetic/curves/src/derive/curve.rs:4-141).
// An affine addition needs one field inversion; shared by a batch through Montgomery's trick it costs 5M + 1S = 788 MAD32
// instead of the XYZZ mixed addition's 1,232 — the accumulation is multiplier-bound, so that is the lever. The shape that fits
// the B200: every thread runs AFF_K independent STREAMS, each a sub-chunk of 2^seg_log consecutive entries of the bucket-sorted
// list with its own affine accumulator in shared memory (64 B per stream, thread-private: no barriers). One step adds the next
// entry of every stream: forward pass (denominators x2 - x1 and their running product, the partial products parked in an
// L2-resident scratch), ONE inversion per thread per step — the branch-free safegcd of fp.cuh, ALU work that overlaps the other
// warps' multiplier work — and a backward pass that peels the inverses off and finishes the additions. No data leaves the
// thread, HBM traffic stays the 64 B gather per entry (the backward pass re-reads the points from L2).
// Exceptional cases as the reference (derive/curve.rs:866-871 / batch_add's own branches): identity base, empty accumulator,
// P + P (the doubling's denominator 2y joins the batch), P + (-P). The first entry of every bucket carries AFF_FIRST (set by
// msm_mark_first_kernel), so a stream sees bucket boundaries without tracking offsets. Outputs are AFFINE (identity = zeros):
// complete buckets, and per-sub-chunk head / tail partials that msm_merge_affine_kernel adds up into the XYZZ bucket array.
constexpr int AFF_THREADS = 128;
constexpr uint32_t AFF_FIRST = 0x80000000u;

__global__ void __launch_bounds__(256) msm_mark_first_kernel(const uint32_t* __restrict__ offs, MsmShape s, uint32_t* __restrict__ sorted) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)s.nsets * s.nb) return;
    const uint32_t w = (uint32_t)(gid / s.nb), d = (uint32_t)(gid % s.nb) + 1;
    const uint32_t* o = offs + (size_t)w * s.stride;
    const uint32_t a = o[d];
    if (o[d + 1] > a) sorted[(size_t)w * s.list_cap + a] |= AFF_FIRST;  // one writer per word
}

// accumulator storage of the streams: shared memory ([stream][quarter][thread], conflict-free 128-bit accesses) or an
// L2-resident global scratch ([stream][quarter][global thread]: coalesced), which lifts the shared-memory limit on the
// number of resident warps
template <bool SMEM>
struct AffAcc {
    uint4* base;
    size_t stride;  // threads per (stream, quarter) plane
    size_t me;      // this thread's slot in a plane
    __device__ __forceinline__ uint4* at(int k, int c) const { return base + ((size_t)(k * 4 + c) * stride + me); }
    __device__ __forceinline__ Fq x(int k) const { return fq_of(*at(k, 0), *at(k, 1)); }
    __device__ __forceinline__ Fq y(int k) const { return fq_of(*at(k, 2), *at(k, 3)); }
    __device__ __forceinline__ void set(int k, const Fq& X, const Fq& Y) const {
        *at(k, 0) = make_uint4(X.l[0], X.l[1], X.l[2], X.l[3]);
        *at(k, 1) = make_uint4(X.l[4], X.l[5], X.l[6], X.l[7]);
        *at(k, 2) = make_uint4(Y.l[0], Y.l[1], Y.l[2], Y.l[3]);
        *at(k, 3) = make_uint4(Y.l[4], Y.l[5], Y.l[6], Y.l[7]);
    }
    static __device__ __forceinline__ Fq fq_of(const uint4& a, const uint4& b) {
        Fq r;
        r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
        return r;
    }
};
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// what the forward pass decided for a stream's entry (2 bits per stream in a register)
enum : uint32_t { AFF_SKIP = 0, AFF_LOAD = 1, AFF_ADD = 2, AFF_DBL = 3 };

// one non-inlined copy of the field multiplication for the two passes (the inlined passes would be ~20 KB of code each)
__device__ __noinline__ Fq aff_mul(Fq a, Fq b) { return fp_mul<FqP>(a, b); }
__device__ __noinline__ Fq aff_inv(Fq a) { return fp_inv_safegcd<FqP>(a); }

template <int AFF_K, bool SMEM>
__global__ void __launch_bounds__(AFF_THREADS, SMEM ? 2 : 4) msm_accumulate_affine_kernel(const uint4* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                                                               const uint32_t* __restrict__ offs, const uint32_t* __restrict__ nxt,
                                                                               MsmShape s, int seg_log, uint32_t cpw, uint32_t tpw,
                                                                               uint4* __restrict__ baff, uint4* __restrict__ head,
                                                                               uint4* __restrict__ tail, uint4* __restrict__ prefix,
                                                                               uint32_t* __restrict__ cur, uint4* __restrict__ acc_scratch) {
    extern __shared__ uint4 aff_smem[];
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)s.nsets * tpw;
    if (gid >= nthreads) return;
    const AffAcc<SMEM> acc{SMEM ? aff_smem : acc_scratch, SMEM ? (size_t)AFF_THREADS : nthreads, SMEM ? (size_t)threadIdx.x : gid};
    const uint32_t w = (uint32_t)(gid / tpw), tj = (uint32_t)(gid % tpw);
    const uint32_t* o = offs + (size_t)w * s.stride;
    const uint32_t* nx = nxt + (size_t)w * s.stride;
    const uint32_t* lst = sorted + (size_t)w * s.list_cap;
    const uint32_t total = o[s.nb + 1];
    const uint32_t seg = 1u << seg_log;
    const uint32_t j0 = tj * AFF_K;                     // first sub-chunk of this thread inside set w
    if (((size_t)j0 << seg_log) >= total) return;
    // streams that hold entries: sub-chunk j0 + k starts below `total`
    int nact = 0;
#pragma unroll 1
    for (int k = 0; k < AFF_K; k++) {
        acc.set(k, Fq::zero(), Fq::zero());
        if ((((size_t)j0 + k) << seg_log) < total) nact = k + 1;
    }
    // current bucket of every stream: the largest d with o[d] <= start (K binary searches side by side)
    {
        uint32_t lo[AFF_K], hi[AFF_K];
#pragma unroll
        for (int k = 0; k < AFF_K; k++) { lo[k] = 1; hi[k] = s.nb; }
        for (uint32_t span = s.nb; span > 1; span = (span + 1) >> 1) {
#pragma unroll
            for (int k = 0; k < AFF_K; k++) {
                if (k < nact && lo[k] < hi[k]) {
                    const uint32_t mid = (lo[k] + hi[k] + 1) >> 1;
                    if (o[mid] <= ((j0 + k) << seg_log)) lo[k] = mid; else hi[k] = mid - 1;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < AFF_K; k++)
            if (k < nact) cur[(size_t)k * nthreads + gid] = lo[k];
    }
    uint32_t first_mask = (1u << AFF_K) - 1u;  // stream has not crossed a bucket boundary yet: its accumulator is the head partial
    uint32_t full_mask = 0;                     // stream's accumulator holds a point
    uint4* const pre = prefix + gid * 2;        // stream k's parked product at pre[k * nthreads * 2]
#pragma unroll 1
    for (uint32_t step = 0; step < seg; step++) {
        // ---- forward: denominators and their running product --------------------------------------------------------------
        Fq run = Fq::one();
        uint64_t kinds = 0;
#pragma unroll 1
        for (int k = 0; k < nact; k++) {
            const uint32_t pos = ((j0 + k) << seg_log) + step;
            if (pos >= total) continue;
            const uint32_t e = __ldg(lst + pos);
            if ((e & AFF_FIRST) && step != 0) {
                // the stream's bucket ended with the previous entry
                const size_t q = (size_t)w * cpw + j0 + k;
                uint32_t* cu = cur + (size_t)k * nthreads + gid;
                const uint32_t d = *cu;
                uint4* dst = (first_mask >> k) & 1u ? head + q * 4 : baff + ((size_t)w * s.nb + (d - 1)) * 4;
                dst[0] = *acc.at(k, 0); dst[1] = *acc.at(k, 1); dst[2] = *acc.at(k, 2); dst[3] = *acc.at(k, 3);
                first_mask &= ~(1u << k);
                full_mask &= ~(1u << k);
                acc.set(k, Fq::zero(), Fq::zero());
                *cu = nx[d + 1];  // next non-empty bucket: the one this entry opens
            }
            const uint4* bp = bases + (size_t)((e & ~AFF_FIRST) >> 1) * 4;
            if (step + 1 < seg && pos + 1 < total) {  // the next step's gather: HBM -> L2 while this step computes
                const uint4* np = bases + (size_t)((__ldg(lst + pos + 1) & ~AFF_FIRST) >> 1) * 4;
                prefetch_l2(np);
                prefetch_l2(np + 2);
            }
            const Fq x2 = ldg_fq(bp);
            Fq y2 = ldg_fq(bp + 2);
            if (x2.is_zero() && y2.is_zero()) continue;  // identity base contributes nothing: reference curve.rs:857-858
            if (!((full_mask >> k) & 1u)) {              // empty accumulator: the backward pass just loads the point
                kinds |= (uint64_t)AFF_LOAD << (2 * k);
                continue;
            }
            uint32_t kind = AFF_ADD;
            Fq den = fp_sub<FqP>(x2, acc.x(k));
            if (den.is_zero()) {
                // same x: the same point (double: the denominator is 2y) or opposite points (the sum is the identity)
                if (e & 1u) y2 = fp_neg<FqP>(y2);
                const Fq y1 = acc.y(k);
                den = fp_dbl<FqP>(y1);
                if (y2 == y1 && !den.is_zero()) kind = AFF_DBL;
                else {
                    full_mask &= ~(1u << k);
                    acc.set(k, Fq::zero(), Fq::zero());
                    continue;
                }
            }
            kinds |= (uint64_t)kind << (2 * k);
            st_fq(pre + (size_t)k * nthreads * 2, run);
            run = aff_mul(run, den);
        }
        if (kinds == 0) continue;
        // ---- one inversion for the whole step --------------------------------------------------------------------------------
        Fq inv = (kinds & 0xaaaaaaaaaaaaaaaaull) ? aff_inv(run) : run;  // no ADD / DBL in this step: nothing to invert
        // ---- backward: peel the inverses off, finish the additions -------------------------------------------------------------
#pragma unroll 1
        for (int k = nact - 1; k >= 0; k--) {
            const uint32_t kind = (uint32_t)(kinds >> (2 * k)) & 3u;
            if (kind == AFF_SKIP) continue;
            const uint32_t pos = ((j0 + k) << seg_log) + step;
            const uint32_t e = __ldg(lst + pos);
            const uint4* bp = bases + (size_t)((e & ~AFF_FIRST) >> 1) * 4;
            Fq x2 = ldg_fq(bp), y2 = ldg_fq(bp + 2);
            if (e & 1u) y2 = fp_neg<FqP>(y2);
            if (kind == AFF_LOAD) {
                acc.set(k, x2, y2);
                full_mask |= 1u << k;
                continue;
            }
            const Fq x1 = acc.x(k), y1 = acc.y(k);
            const Fq den = kind == AFF_ADD ? fp_sub<FqP>(x2, x1) : fp_dbl<FqP>(y1);
            const Fq dinv = aff_mul(inv, ld_fq(pre + (size_t)k * nthreads * 2));
            inv = aff_mul(inv, den);
            Fq num;
            if (kind == AFF_ADD) num = fp_sub<FqP>(y2, y1);
            else { const Fq xx = fp_sqr<FqP>(x1); num = fp_add<FqP>(fp_dbl<FqP>(xx), xx); }  // 3 x^2 (a = 0)
            const Fq lam = aff_mul(num, dinv);
            const Fq x3 = fp_sub<FqP>(fp_sub<FqP>(fp_sqr<FqP>(lam), x1), x2);  // x2 == x1 in the doubling
            const Fq y3 = fp_sub<FqP>(aff_mul(lam, fp_sub<FqP>(x1, x3)), y1);
            acc.set(k, x3, y3);
        }
    }
#pragma unroll 1
    for (int k = 0; k < nact; k++) {
        const size_t q = (size_t)w * cpw + j0 + k;
        uint4* dst = ((first_mask >> k) & 1u ? head : tail) + q * 4;
        dst[0] = *acc.at(k, 0); dst[1] = *acc.at(k, 1); dst[2] = *acc.at(k, 2); dst[3] = *acc.at(k, 3);
    }
}

// buckets[0][i] += buckets[p][i], p = 1..nparts-1: the bucket arrays of a pipelined MSM's parts are folded by one thread
// per bucket (a throughput kernel) before the latency-sized reduction, which then reads a single array
__global__ void __launch_bounds__(128) msm_fold_parts_kernel(uint4* __restrict__ buckets, size_t nbuckets, int nparts) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbuckets) return;
    G1Xyzz acc = ld_xyzz(buckets + i * 8);
    for (int p = 1; p < nparts; p++) {
        G1Xyzz b = ld_xyzz(buckets + ((size_t)p * nbuckets + i) * 8);
        g1_add(acc, b);
    }
    st_xyzz(buckets + i * 8, acc);
}

// the kernels from here to the precompute kernel are latency-bound chains of XYZZ operations: compact code (ec.cuh FqCall)
// (round 2 re-check: inlined multiplications INSIDE the one non-inlined point addition / doubling of each kernel, 65 KB of code, are
// slower again — bucket reduction 0.19 -> 0.32 ms at 2^16, 0.64 -> 0.76 ms at 2^24 — so the call form stays)
typedef FqCall TailMul;
// one copy of each point operation per kernel as well (a g1_add is still ~1.2k instructions around its 13 multiplier calls)
__device__ __noinline__ G1Xyzz tail_add_fn(G1Xyzz a, G1Xyzz b) { g1_add<TailMul>(a, b); return a; }
__device__ __noinline__ G1Xyzz tail_double_fn(G1Xyzz a) { return g1_double<TailMul>(a); }
__device__ __forceinline__ void tail_add(G1Xyzz& acc, const G1Xyzz& b) { acc = tail_add_fn(acc, b); }
__device__ __forceinline__ G1Xyzz tail_double(const G1Xyzz& a) { return tail_double_fn(a); }

// (4b) one thread per bucket: add up the head/tail partials of the chunks the bucket overlaps. Buckets that span more
// than MERGE_LONG chunks (heavily repeated digits) are queued for msm_merge_big_kernel.
constexpr uint32_t MERGE_LONG = 32;
__global__ void __launch_bounds__(128) msm_merge_kernel(const uint32_t* __restrict__ offs, MsmShape s, int seg_log, uint32_t cpw,
                                                        uint4* __restrict__ buckets, const uint4* __restrict__ head,
                                                        const uint4* __restrict__ tail, uint32_t* __restrict__ big_count,
                                                        uint2* __restrict__ big_list, uint32_t big_cap) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)s.nsets * s.nb) return;
    uint32_t w = (uint32_t)(gid / s.nb), d = (uint32_t)(gid % s.nb) + 1;
    const uint32_t* o = offs + (size_t)w * s.stride;
    const uint32_t total = o[s.nb + 1];
    const uint32_t start = o[d], end = o[d + 1];
    if (start == end) return;  // empty bucket: stays identity (buckets are zero-initialised, ZZ = 0)
    const uint32_t kf = start >> seg_log, kl = (end - 1) >> seg_log;
    const uint32_t seg = 1u << seg_log;
    if (kf == kl) {
        uint32_t cs = kf << seg_log, ce = min(cs + seg, total);
        bool isfirst = start <= cs, islast = end >= ce;
        if (!isfirst && !islast) return;  // complete bucket, already stored by the chunk thread
        const uint4* src = (isfirst ? head : tail) + ((size_t)w * cpw + kf) * 8;
        G1Xyzz p = ld_xyzz(src);
        st_xyzz(buckets + gid * 8, p);
        return;
    }
    if (kl - kf > MERGE_LONG) {
        uint32_t slot = atomicAdd(big_count, 1u);
        if (slot < big_cap) big_list[slot] = make_uint2(w, d);
        return;
    }
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t k = kf; k <= kl; k++) {
        bool isfirst = start <= (k << seg_log);
        G1Xyzz p = ld_xyzz((isfirst ? head : tail) + ((size_t)w * cpw + k) * 8);
        tail_add(acc, p);
    }
    st_xyzz(buckets + gid * 8, acc);
}

// shared-memory tree of XYZZ adds over NT threads (power of two); result valid in thread 0
template <int NT>
__device__ __forceinline__ G1Xyzz block_sum(G1Xyzz acc, uint4* sm) {
    const int tid = threadIdx.x;
    st_xyzz(sm + tid * 8, acc);
    __syncthreads();
#pragma unroll 1
    for (int half = NT / 2; half >= 1; half >>= 1) {
        if (tid < half) {
            G1Xyzz a = ld_xyzz(sm + tid * 8), b2 = ld_xyzz(sm + (tid + half) * 8);
            tail_add(a, b2);
            st_xyzz(sm + tid * 8, a);
        }
        __syncthreads();
    }
    return ld_xyzz(sm);
}
__device__ __forceinline__ G1Xyzz block_sum_128(G1Xyzz acc, uint4* sm) { return block_sum<128>(acc, sm); }

// (4c) one CTA per queued long bucket: threads stride over its chunk partials, then a shared-memory tree of XYZZ adds
__global__ void __launch_bounds__(128) msm_merge_big_kernel(const uint32_t* __restrict__ offs, MsmShape s, int seg_log, uint32_t cpw,
                                                            uint4* __restrict__ buckets, const uint4* __restrict__ head,
                                                            const uint4* __restrict__ tail, const uint32_t* __restrict__ big_count,
                                                            const uint2* __restrict__ big_list) {
    __shared__ uint4 sm[128 * 8];
    if (blockIdx.x >= *big_count) return;
    const uint2 wd = big_list[blockIdx.x];
    const uint32_t w = wd.x, d = wd.y;
    const uint32_t* o = offs + (size_t)w * s.stride;
    const uint32_t start = o[d], end = o[d + 1];
    const uint32_t kf = start >> seg_log, kl = (end - 1) >> seg_log;
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t k = kf + threadIdx.x; k <= kl; k += 128) {
        bool isfirst = start <= (k << seg_log);
        G1Xyzz p = ld_xyzz((isfirst ? head : tail) + ((size_t)w * cpw + k) * 8);
        tail_add(acc, p);
    }
    G1Xyzz r = block_sum_128(acc, sm);
    if (threadIdx.x == 0) st_xyzz(buckets + ((size_t)w * s.nb + (d - 1)) * 8, r);
}

// (4b') the same two kernels for the batched-affine accumulation: partials and complete buckets are affine points (zeros =
// identity); the XYZZ bucket array the reduction reads is produced here.
__device__ __forceinline__ void tail_madd_affine(G1Xyzz& acc, const uint4* p) {
    const Fq x = ld_fq(p), y = ld_fq(p + 2);
    if (x.is_zero() && y.is_zero()) return;
    g1_madd<TailMul>(acc, x, y);
}
__global__ void __launch_bounds__(128) msm_merge_affine_kernel(const uint32_t* __restrict__ offs, MsmShape s, int seg_log, uint32_t cpw,
                                                               uint4* __restrict__ buckets, const uint4* __restrict__ baff,
                                                               const uint4* __restrict__ head, const uint4* __restrict__ tail,
                                                               uint32_t* __restrict__ big_count, uint2* __restrict__ big_list, uint32_t big_cap) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)s.nsets * s.nb) return;
    uint32_t w = (uint32_t)(gid / s.nb), d = (uint32_t)(gid % s.nb) + 1;
    const uint32_t* o = offs + (size_t)w * s.stride;
    const uint32_t total = o[s.nb + 1];
    const uint32_t start = o[d], end = o[d + 1];
    if (start == end) return;  // empty bucket: stays identity (the XYZZ array is zero-initialised)
    const uint32_t kf = start >> seg_log, kl = (end - 1) >> seg_log;
    const uint32_t seg = 1u << seg_log;
    G1Xyzz acc = G1Xyzz::identity();
    if (kf == kl) {
        uint32_t cs = kf << seg_log, ce = min(cs + seg, total);
        bool isfirst = start <= cs, islast = end >= ce;
        const uint4* src = (!isfirst && !islast) ? baff + gid * 4 : (isfirst ? head : tail) + ((size_t)w * cpw + kf) * 4;
        const Fq x = ld_fq(src), y = ld_fq(src + 2);
        if (!(x.is_zero() && y.is_zero())) { acc.x = x; acc.y = y; acc.zz = Fq::one(); acc.zzz = Fq::one(); }
        st_xyzz(buckets + gid * 8, acc);
        return;
    }
    if (kl - kf > MERGE_LONG) {
        uint32_t slot = atomicAdd(big_count, 1u);
        if (slot < big_cap) big_list[slot] = make_uint2(w, d);
        return;
    }
    for (uint32_t k = kf; k <= kl; k++) {
        bool isfirst = start <= (k << seg_log);
        tail_madd_affine(acc, (isfirst ? head : tail) + ((size_t)w * cpw + k) * 4);
    }
    st_xyzz(buckets + gid * 8, acc);
}
constexpr int MERGE_BIG_AFF_THREADS = 512;
__global__ void __launch_bounds__(MERGE_BIG_AFF_THREADS) msm_merge_big_affine_kernel(const uint32_t* __restrict__ offs, MsmShape s, int seg_log,
                                                                                     uint32_t cpw, uint4* __restrict__ buckets,
                                                                                     const uint4* __restrict__ head, const uint4* __restrict__ tail,
                                                                                     const uint32_t* __restrict__ big_count,
                                                                                     const uint2* __restrict__ big_list) {
    extern __shared__ uint4 big_sm[];
    if (blockIdx.x >= *big_count) return;
    const uint2 wd = big_list[blockIdx.x];
    const uint32_t w = wd.x, d = wd.y;
    const uint32_t* o = offs + (size_t)w * s.stride;
    const uint32_t start = o[d], end = o[d + 1];
    const uint32_t kf = start >> seg_log, kl = (end - 1) >> seg_log;
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t k = kf + threadIdx.x; k <= kl; k += MERGE_BIG_AFF_THREADS) {
        bool isfirst = start <= (k << seg_log);
        tail_madd_affine(acc, (isfirst ? head : tail) + ((size_t)w * cpw + k) * 4);
    }
    G1Xyzz r = block_sum<MERGE_BIG_AFF_THREADS>(acc, big_sm);
    if (threadIdx.x == 0) st_xyzz(buckets + ((size_t)w * s.nb + (d - 1)) * 8, r);
}

// (5) per-set weighted bucket sum, chunked: thread t of set w owns bucket ids [t*CH + 1, (t+1)*CH]; the CTA's RED_CTA
// results are tree-added in shared memory and ONE partial per CTA is stored (tpw is a multiple of RED_CTA, so a CTA never
// straddles two sets). Sized for latency, not occupancy: about one warp per SM sub-partition (the whole phase is a chain
// of dependent XYZZ operations per thread; more threads with shorter chains only add multiply-by-t overhead).
constexpr int RED_CTA = 32;
__global__ void __launch_bounds__(RED_CTA) msm_reduce_kernel(const uint4* __restrict__ buckets, MsmShape s, uint32_t tpw, uint32_t ch,
                                                             int nparts, size_t part_stride, uint4* __restrict__ partials) {
    __shared__ uint4 sm[RED_CTA * 8];
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    G1Xyzz acc = G1Xyzz::identity();
    const uint32_t w = (uint32_t)(gid / tpw), t = (uint32_t)(gid % tpw);
    if (w < (uint32_t)s.nsets && (size_t)t * ch < s.nb) {  // tpw is padded to a multiple of RED_CTA for tiny bucket sets
        const uint4* b = buckets + ((size_t)w * s.nb + (size_t)t * ch) * 8;
        G1Xyzz running = G1Xyzz::identity();
        for (int j = (int)ch - 1; j >= 0; j--) {  // summation by parts, reference arithmetic.rs:95-99
            G1Xyzz bj = ld_xyzz(b + (size_t)j * 8);
            for (int p = 1; p < nparts; p++) {  // parts of a pipelined host-pointer MSM
                G1Xyzz bp = ld_xyzz(b + (size_t)p * part_stride + (size_t)j * 8);
                tail_add(bj, bp);
            }
            tail_add(running, bj);
            tail_add(acc, running);
        }
        // acc = sum (j+1) * B ; add (t*ch) * running
        if (t != 0 && !running.is_identity()) {
            G1Xyzz m = G1Xyzz::identity();
            for (int bit = 31 - __clz(t); bit >= 0; bit--) {
                m = tail_double(m);
                if ((t >> bit) & 1u) tail_add(m, running);
            }
            for (uint32_t x = ch; x > 1; x >>= 1) m = tail_double(m);  // ch is a power of two
            tail_add(acc, m);
        }
    }
    G1Xyzz r = block_sum<RED_CTA>(acc, sm);
    if (threadIdx.x == 0) st_xyzz(partials + (size_t)blockIdx.x * 8, r);
}

// (5b) segmented tree sum: set w has `count` XYZZ inputs; CTA (w, j) adds inputs [j*2048, (j+1)*2048) into out[w][j]
__global__ void __launch_bounds__(128) xyzz_sum_kernel(const uint4* __restrict__ in, uint32_t count, uint32_t out_per_set,
                                                       uint4* __restrict__ out) {
    __shared__ uint4 sm[128 * 8];
    const uint32_t w = blockIdx.x / out_per_set, j = blockIdx.x % out_per_set;
    const uint32_t lo = j * 2048u, hi = min(lo + 2048u, count);
    G1Xyzz acc = G1Xyzz::identity();
    for (uint32_t k = lo + threadIdx.x; k < hi; k += 128) {
        G1Xyzz p = ld_xyzz(in + ((size_t)w * count + k) * 8);
        tail_add(acc, p);
    }
    G1Xyzz r = block_sum_128(acc, sm);
    if (threadIdx.x == 0) st_xyzz(out + ((size_t)w * out_per_set + j) * 8, r);
}

// (6) Horner over the bucket sets + affine normalisation: out = 64 B x||y, then uint32 is_identity
__global__ void msm_final_kernel(const uint4* __restrict__ wins, MsmShape s, uint4* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1Xyzz acc = G1Xyzz::identity();
    for (int w = s.nsets - 1; w >= 0; w--) {
        if (!acc.is_identity())
            for (int k = 0; k < s.c; k++) acc = tail_double(acc);  // reference arithmetic.rs:47-49
        G1Xyzz ww = ld_xyzz(wins + (size_t)w * 8);
        tail_add(acc, ww);
    }
    G1Affine a = g1_to_affine_lowlat<TailMul>(acc);
    st_fq(out, a.x);
    st_fq(out + 2, a.y);
    uint32_t inf = acc.is_identity() ? 1u : 0u;
    out[4] = make_uint4(inf, 0, 0, 0);
}

// batched single-layout MSMs: bucket set b is MSM b; no Horner, one normalisation per block
__global__ void msm_final_batch_kernel(const uint4* __restrict__ wins, uint4* __restrict__ out) {
    if (threadIdx.x != 0) return;
    const uint32_t b = blockIdx.x;
    G1Xyzz acc = ld_xyzz(wins + (size_t)b * 8);
    G1Affine a = g1_to_affine_lowlat<TailMul>(acc);
    st_fq(out + (size_t)b * 5, a.x);
    st_fq(out + (size_t)b * 5 + 2, a.y);
    out[(size_t)b * 5 + 4] = make_uint4(acc.is_identity() ? 1u : 0u, 0, 0, 0);
}

// sum of n affine points (multi-GPU partial fold). Single CTA; n is tiny (number of GPUs) but any n works.
__global__ void __launch_bounds__(128) g1_sum_affine_kernel(const uint4* __restrict__ pts, size_t n, uint4* __restrict__ out) {
    __shared__ uint4 sm[128 * 8];
    G1Xyzz acc = G1Xyzz::identity();
    for (size_t j = threadIdx.x; j < n; j += 128) {
        Fq x = ld_fq(pts + j * 4), y = ld_fq(pts + j * 4 + 2);
        if (x.is_zero() && y.is_zero()) continue;
        g1_madd<TailMul>(acc, x, y);
    }
    G1Xyzz r = block_sum_128(acc, sm);
    if (threadIdx.x == 0) {
        G1Affine a = g1_to_affine_lowlat<TailMul>(r);
        st_fq(out, a.x);
        st_fq(out + 2, a.y);
        out[4] = make_uint4(r.is_identity() ? 1u : 0u, 0, 0, 0);
    }
}

// ---- precomputed table: row w holds 2^(c w) P_i. One launch per row: each thread turns PRE_RUN consecutive points of
// row w-1 into row w with c doublings in XYZZ and ONE shared inversion (batch_normalize, derive/curve.rs:362-397).
constexpr int PRE_RUN = 32;
__global__ void __launch_bounds__(128) msm_precompute_row_kernel(const uint4* __restrict__ prev, uint4* __restrict__ next, size_t n, int c,
                                                                 uint4* __restrict__ tmp) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t p0 = t * PRE_RUN;
    if (p0 >= n) return;
    size_t cnt = (n - p0 < (size_t)PRE_RUN) ? (n - p0) : (size_t)PRE_RUN;
    Fq prod = Fq::one();
    for (size_t j = 0; j < cnt; j++) {
        const uint4* src = prev + (p0 + j) * 4;
        Fq x = ld_fq(src), y = ld_fq(src + 2);
        G1Xyzz P = G1Xyzz::identity();
        if (!(x.is_zero() && y.is_zero())) {
            P = g1_double_affine(x, y);
            for (int k = 1; k < c; k++) P = g1_double(P);
        }
        uint4* slot = tmp + (p0 + j) * 10;
        st_xyzz(slot, P);
        st_fq(slot + 8, prod);
        if (!P.is_identity()) prod = fp_mul<FqP>(prod, fp_mul<FqP>(P.zz, P.zzz));
    }
    Fq inv = fp_inv<FqP>(prod);
    for (size_t j = cnt; j-- > 0;) {
        const uint4* slot = tmp + (p0 + j) * 10;
        G1Xyzz P = ld_xyzz(slot);
        Fq pre = ld_fq(slot + 8);
        Fq ax = Fq::zero(), ay = Fq::zero();
        if (!P.is_identity()) {
            Fq zi = fp_mul<FqP>(inv, pre);
            inv = fp_mul<FqP>(inv, fp_mul<FqP>(P.zz, P.zzz));
            ax = fp_mul<FqP>(P.x, fp_mul<FqP>(zi, P.zzz));
            ay = fp_mul<FqP>(P.y, fp_mul<FqP>(zi, P.zz));
        }
        uint4* dst = next + (p0 + j) * 4;
        st_fq(dst, ax);
        st_fq(dst + 2, ay);
    }
}

// ---- Curve::batch_normalize (arithmetic/curves/src/derive/curve.rs:362-397): Jacobian (x, y, z) -> affine (x / z^2, y / z^3),
// identities (z = 0) -> (0, 0). MSMKZG::eval (poly/kzg/msm.rs:65-70) normalises its projective bases this way before
// best_multiexp. One thread per run of NORM_RUN points: prefix products of the z's, ONE inversion (safegcd), back substitution.
constexpr int NORM_RUN = 16;
__global__ void __launch_bounds__(128) g1_batch_normalize_kernel(const uint4* __restrict__ jac, size_t n, uint4* __restrict__ aff) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t p0 = t * NORM_RUN;
    if (p0 >= n) return;
    const int cnt = (int)((n - p0 < (size_t)NORM_RUN) ? (n - p0) : (size_t)NORM_RUN);
    Fq pre[NORM_RUN];
    Fq acc = Fq::one();
#pragma unroll 1
    for (int j = 0; j < cnt; j++) {
        pre[j] = acc;
        const Fq z = ld_fq(jac + (p0 + j) * 6 + 4);
        if (!z.is_zero()) acc = fp_mul<FqP>(acc, z);
    }
    acc = fp_inv_safegcd<FqP>(acc);
#pragma unroll 1
    for (int j = cnt - 1; j >= 0; j--) {
        const uint4* src = jac + (p0 + j) * 6;
        const Fq z = ld_fq(src + 4);
        Fq ax = Fq::zero(), ay = Fq::zero();
        if (!z.is_zero()) {
            const Fq zi = fp_mul<FqP>(pre[j], acc);  // 1 / z
            acc = fp_mul<FqP>(acc, z);
            const Fq zi2 = fp_sqr<FqP>(zi);
            ax = fp_mul<FqP>(ld_fq(src), zi2);
            ay = fp_mul<FqP>(ld_fq(src + 2), fp_mul<FqP>(zi2, zi));
        }
        st_fq(aff + (p0 + j) * 4, ax);
        st_fq(aff + (p0 + j) * 4 + 2, ay);
    }
}

// -------------------------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------------------------
static PerDevice<Scratch> g_hist, g_sorted, g_buckets, g_partials, g_chunks, g_pre_tmp, g_aff;

// Second stream (high priority) for the sort phases (count / scan / scatter) of part p+1 of a large MSM, which overlap the
// bucket accumulation of part p on the main stream: the sort is atomics/latency bound, the accumulation multiplier bound.
constexpr int MSM_MAX_PARTS = 8;
struct ProfSpan { int phase; cudaEvent_t a, b; };
struct MsmDev {  // per device slot (context.h): streams, events and profiling spans of the MSMs running on that device
    cudaStream_t sort_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_sorted[MSM_MAX_PARTS];
    std::vector<ProfSpan> spans;
    size_t spans_used = 0;
    bool aff_attr_set = false;
};
static PerDevice<MsmDev> g_dev;
#define g_sort_stream (g_dev->sort_stream)
#define g_ev_start (g_dev->ev_start)
#define g_ev_sorted (g_dev->ev_sorted)
#define g_spans (g_dev->spans)
#define g_spans_used (g_dev->spans_used)
static int ensure_sort_stream() {
    if (g_sort_stream) return 0;
    int lo = 0, hi = 0;
    CQB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CQB_CUDA(cudaStreamCreateWithPriority(&g_sort_stream, cudaStreamNonBlocking, hi));
    CQB_CUDA(cudaEventCreateWithFlags(&g_ev_start, cudaEventDisableTiming));
    for (auto& e : g_ev_sorted) CQB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return 0;
}

// optional per-phase device timing: spans of (phase id, start event, end event) recorded on whichever stream the phase
// runs on; msm_phase_ms sums the spans of each phase (a pipelined MSM has one span per phase and part). No host sync
// until the times are read.
static bool g_prof = false;
void msm_set_profiling(bool on) { g_prof = on; }
static int prof_begin(int phase, cudaStream_t st) {
    if (!g_prof) return -1;
    if (g_spans_used == g_spans.size()) {
        ProfSpan sp;
        sp.phase = 0;
        cudaEventCreate(&sp.a);
        cudaEventCreate(&sp.b);
        g_spans.push_back(sp);
    }
    g_spans[g_spans_used].phase = phase;
    cudaEventRecord(g_spans[g_spans_used].a, st);
    return (int)g_spans_used++;
}
static void prof_end(int h, cudaStream_t st) {
    if (h >= 0) cudaEventRecord(g_spans[h].b, st);
}
// ms[0..7] = count, scan, scatter, accumulate, merge, reduce, sum, final of the most recent MSM (summed over its parts)
int msm_phase_ms(float* ms, int cap) {
    if (!g_prof || g_spans_used == 0) return 0;
    int k = std::min(cap, 8);
    for (int i = 0; i < k; i++) ms[i] = 0.f;
    for (size_t i = 0; i < g_spans_used; i++) {
        float t = 0.f;
        cudaEventSynchronize(g_spans[i].b);
        cudaEventElapsedTime(&t, g_spans[i].a, g_spans[i].b);
        if (g_spans[i].phase < k) ms[g_spans[i].phase] += t;
    }
    return k;
}

void msm_release_all() {
    g_hist->release();
    g_sorted->release();
    g_buckets->release();
    g_partials->release();
    g_chunks->release();
    g_pre_tmp->release();
    g_aff->release();
    for (auto& sp : g_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    g_spans.clear();
    g_spans_used = 0;
    if (g_sort_stream) {
        cudaStreamDestroy(g_sort_stream);
        g_sort_stream = nullptr;
        cudaEventDestroy(g_ev_start);
        for (auto& e : g_ev_sorted) cudaEventDestroy(e);
    }
}

static int ceil_log2(size_t n) {
    int l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

int msm_windows_for(int c) { return 254 / c + 1; }  // scalars < r < 2^254; the extra window absorbs the signed-digit carry

// window bits for the single-set layout: minimise nwin(c) * n * 10 (bucket additions) + 2^(c-1) * 40 (bucket reduction)
int msm_precompute_window_bits(size_t n) {
    if (g_forced_c >= 8) return std::min(g_forced_c, 23);
    double best = 1e300;
    int best_c = 12;
    for (int c = 10; c <= 20; c++) {  // beyond 2^19 buckets the scatter's open write streams thrash L2
        // measured (tools/sweep_msm.py, r02): 0.149 ns per bucket addition, bucket reduction 0.18 ms + 0.9 ns per bucket (2^21 points:
        // c = 20 5.78 ms, c = 17 5.99 ms — the 130 of round 1 kept c = 17 there and cost the 8-GPU run two extra windows)
        double cost = msm_windows_for(c) * (double)n * 10.0 + (double)((size_t)1 << (c - 1)) * 60.0;
        if (cost < best) { best = cost; best_c = c; }
    }
    return best_c;
}

// builds the table: row 0 = the bases themselves (copied), row w = 2^c * row (w-1); d_table has nwin * n * 64 bytes
int msm_precompute_table(const void* d_bases, size_t n, int c, void* d_table) {
    cudaStream_t st = ctx().stream;
    int nwin = msm_windows_for(c);
    CQB_CUDA(cudaMemcpyAsync(d_table, d_bases, n * 64, cudaMemcpyDeviceToDevice, st));
    const size_t CHUNK = (size_t)1 << 22;
    CQB_TRY(g_pre_tmp->ensure(std::min(n, CHUNK) * 160));
    for (int w = 1; w < nwin; w++) {
        const char* prev = (const char*)d_table + (size_t)(w - 1) * n * 64;
        char* next = (char*)d_table + (size_t)w * n * 64;
        for (size_t off = 0; off < n; off += CHUNK) {
            size_t m = std::min(CHUNK, n - off);
            size_t threads = (m + PRE_RUN - 1) / PRE_RUN;
            msm_precompute_row_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>((const uint4*)(prev + off * 64), (uint4*)(next + off * 64), m, c,
                                                                                          g_pre_tmp->as<uint4>());
            CQB_LAUNCHED();
        }
    }
    CQB_CUDA(cudaGetLastError());
    return 0;
}

static MsmShape windowed_shape(size_t n) {
    int c;
    if (g_forced_c > 0) c = std::min(g_forced_c, 16);
    else {
        c = ceil_log2(n) - 4;
        if (c < 4) c = 4;
        if (c > 16) c = 16;
    }
    if (c < 2) c = 2;
    MsmShape s;
    s.c = c;
    s.nwin = msm_windows_for(c);
    s.nsets = s.nwin;
    s.single = 0;
    s.nb = 1u << (c - 1);
    s.stride = s.nb + 2;
    s.table_n = 0;
    s.offset = 0;
    s.list_cap = n;
    return s;
}

// An MSM runs as `nparts` PARTS over contiguous point ranges (1 part unless it is large). Each part has its own histogram
// / offsets / sorted list / bucket array; its SORT phase (count, scan, scatter) may run on the sort stream while the
// ACCUMULATE phase (accumulate, merge) of the previous part runs on the main stream. msm_finish adds the parts' bucket
// arrays while it reduces them. For a host-pointer MSM the sort of part p additionally waits for the H2D copy of part p.
struct PartBuf {
    uint32_t *hist, *offs, *cursor, *tile_sums, *nonempty;
    uint32_t* sorted;
    uint4* buckets;
};
struct PartPlan {
    size_t hist_words, part_hist_words, nbuckets;
    uint32_t ntiles;
};
static int plan_parts(const MsmShape& s, int nparts, PartPlan* pl) {
    pl->hist_words = (size_t)s.nsets * s.stride;
    pl->ntiles = (s.nb + 1 + SCAN_TILE - 1) / SCAN_TILE;
    pl->part_hist_words = pl->hist_words * 3 + 2 * (size_t)pl->ntiles + 8;
    pl->nbuckets = (size_t)s.nsets * s.nb;
    CQB_TRY(g_hist->ensure((size_t)nparts * pl->part_hist_words * 4));
    CQB_TRY(g_sorted->ensure((size_t)nparts * s.nsets * s.list_cap * 4));
    CQB_TRY(g_buckets->ensure((size_t)nparts * pl->nbuckets * 128));
    return 0;
}
static PartBuf part_buf(const MsmShape& s, const PartPlan& pl, int part) {
    PartBuf b;
    b.hist = g_hist->as<uint32_t>() + (size_t)part * pl.part_hist_words;
    b.offs = b.hist + pl.hist_words;
    b.cursor = b.offs + pl.hist_words;
    b.tile_sums = b.cursor + pl.hist_words;
    b.nonempty = b.tile_sums + 2 * (size_t)pl.ntiles;  // one word (of the 8 spare ones): occupied buckets of this part
    b.sorted = g_sorted->as<uint32_t>() + (size_t)part * s.nsets * s.list_cap;
    b.buckets = g_buckets->as<uint4>() + (size_t)part * pl.nbuckets * 8;
    return b;
}

// d_bases: windowed layout -> element 0 of the registered set (pid = offset + i); single-set layout -> the table base.
static int msm_sort_phase(const void* d_scalars, const uint32_t* d_idx, size_t n, const MsmShape& s, const PartPlan& pl, const PartBuf& b,
                          cudaStream_t st) {
    CQB_CUDA(cudaMemsetAsync(b.hist, 0, pl.hist_words * 4, st));
    CQB_CUDA(cudaMemsetAsync(b.nonempty, 0, 4, st));
    int h = prof_begin(0, st);
    unsigned gN = (unsigned)((n + 255) / 256);
    const dim3 gridN(gN, s.single ? (unsigned)s.nsets : 1u);
    msm_count_kernel<<<gridN, 256, 0, st>>>((const uint4*)d_scalars, n, s, b.hist);
    CQB_LAUNCHED();
    prof_end(h, st);
    h = prof_begin(1, st);
    if (s.single && s.nb > 32768) {
        for (int k = 0; k < s.nsets; k++) {  // one tiled scan per bucket set
            const size_t o = (size_t)k * s.stride;
            scan_tile_sums_kernel<<<pl.ntiles, 1024, 0, st>>>(b.hist + o, s.nb, b.tile_sums, pl.ntiles);
            CQB_LAUNCHED();
            scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(b.tile_sums, pl.ntiles);
            CQB_LAUNCHED();
            scan_apply_kernel<<<pl.ntiles, 1024, 0, st>>>(b.hist + o, s.nb, b.tile_sums, pl.ntiles, b.offs + o, b.cursor + o, b.nonempty);
            CQB_LAUNCHED();
        }
    } else {
        msm_scan_kernel<<<s.nsets, 1024, 0, st>>>(b.hist, b.offs, b.cursor, s, b.nonempty);
        CQB_LAUNCHED();
    }
    prof_end(h, st);
    h = prof_begin(2, st);
    // bucket-range passes of the scatter: keep <= 2^17 write streams open at a time
    int range_shift = 31;
    unsigned passes = 1;
    if (s.nb >= (1u << 19) && n >= ((size_t)1 << 23)) passes = 2;  // measured at 2^24 / c = 20: 3.66 ms (1 pass), 3.14 (2), 4.19 (4: each pass re-derives the digits)
    if (passes > 1) {
        int lp = 0;
        while ((1u << lp) < passes) lp++;
        passes = 1u << lp;
        range_shift = std::max(0, (s.c - 1) - lp);
        passes = (s.nb + (1u << range_shift) - 1) >> range_shift;
    }
    const dim3 gridS(gN, s.single ? (unsigned)s.nsets : 1u, passes);
    msm_scatter_kernel<<<gridS, 256, 0, st>>>((const uint4*)d_scalars, d_idx, n, s, b.cursor, b.sorted, range_shift);
    CQB_LAUNCHED();
    prof_end(h, st);
    CQB_CUDA(cudaGetLastError());
    return 0;
}

// bucket accumulation variant: 0 = automatic (= XYZZ: the batched-affine kernel measured slower on B200), 1 = XYZZ mixed additions, 2 = batched affine
static int g_acc_mode = 0;
void msm_set_accumulator(int mode) { g_acc_mode = mode; }
static int g_aff_seg_log = 0;  // experiments: entries per stream = 2^g_aff_seg_log (0 = automatic)
void msm_set_affine_segment(int seg_log) { g_aff_seg_log = seg_log; }

// batched-affine accumulation + merge (kernel comment above); the XYZZ bucket array comes out as msm_acc_phase leaves it
static int g_aff_variant = 0;  // experiments: 0 = accumulators in shared memory, K = 14 | 1 = global scratch, K = 14 | 2 = global scratch, K = 28
void msm_set_affine_variant(int v) { g_aff_variant = v; }
static int msm_acc_phase_affine(const void* d_bases, size_t n, const MsmShape& s, const PartPlan& pl, const PartBuf& b, cudaStream_t st) {
    bool& attr_set = g_dev->aff_attr_set;
    const int K = g_aff_variant == 2 ? 28 : 14;
    const bool in_smem = g_aff_variant == 0;
    const size_t smem = in_smem ? (size_t)K * 64 * AFF_THREADS : 0;
    if (!attr_set) {
        CQB_CUDA(cudaFuncSetAttribute(msm_accumulate_affine_kernel<14, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 14 * 64 * AFF_THREADS));
        CQB_CUDA(cudaFuncSetAttribute(msm_merge_big_affine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MERGE_BIG_AFF_THREADS * 128));
        attr_set = true;
    }
    const size_t entries = (size_t)n * s.nwin;
    // entries per stream: 32, shrinking while the grid would be shorter than ~4 waves of 2 CTAs x 148 SMs
    int seg_log = 5;
    const size_t wave = (size_t)2 * ctx().sm_count * AFF_THREADS * K;
    while (seg_log > 3 && (entries >> seg_log) < 4 * wave) seg_log--;
    if (g_aff_seg_log > 0) seg_log = g_aff_seg_log;
    const size_t list_len = s.single ? entries : n;
    const uint32_t cpw = (uint32_t)((list_len + ((size_t)1 << seg_log) - 1) >> seg_log);  // sub-chunks (streams) per set
    const uint32_t tpw = (cpw + K - 1) / K;                                                // threads per set
    const size_t nsub = (size_t)s.nsets * cpw, nthreads = (size_t)s.nsets * tpw;
    const uint32_t big_cap = (uint32_t)(s.nsets * (cpw / MERGE_LONG + 2));
    // head, tail: 64 B per sub-chunk; baff: 64 B per bucket; parked products: 32 B per stream; current bucket: 4 B per stream
    const size_t head_b = nsub * 64, baff_b = pl.nbuckets * 64, pre_b = nthreads * K * 32, cur_b = (nthreads * K * 4 + 15) / 16 * 16;
    const size_t acc_b = in_smem ? 0 : nthreads * K * 64;
    CQB_TRY(g_aff->ensure(2 * head_b + baff_b + pre_b + acc_b + cur_b + 16 + (size_t)big_cap * 8));
    uint4* head = g_aff->as<uint4>();
    uint4* tail = head + nsub * 4;
    uint4* baff = tail + nsub * 4;
    uint4* prefix = baff + pl.nbuckets * 4;
    uint4* accs = prefix + nthreads * K * 2;
    uint32_t* cur = (uint32_t*)(accs + acc_b / 16);
    uint32_t* big_count = cur + cur_b / 4;
    uint2* big_list = (uint2*)(big_count + 4);
    CQB_CUDA(cudaMemsetAsync(b.buckets, 0, pl.nbuckets * 128, st));
    CQB_CUDA(cudaMemsetAsync(big_count, 0, 16, st));
    int h = prof_begin(3, st);
    msm_mark_first_kernel<<<(unsigned)((pl.nbuckets + 255) / 256), 256, 0, st>>>(b.offs, s, b.sorted);
    CQB_LAUNCHED();
    const unsigned grid = (unsigned)((nthreads + AFF_THREADS - 1) / AFF_THREADS);
#define AFF_LAUNCH(KK, SM)                                                                                                              \
    msm_accumulate_affine_kernel<KK, SM><<<grid, AFF_THREADS, smem, st>>>((const uint4*)d_bases, b.sorted, b.offs, b.hist, s, seg_log, cpw, tpw, \
                                                                         baff, head, tail, prefix, cur, accs)
    if (g_aff_variant == 0) AFF_LAUNCH(14, true);
    else if (g_aff_variant == 1) AFF_LAUNCH(14, false);
    else AFF_LAUNCH(28, false);
#undef AFF_LAUNCH
    CQB_LAUNCHED();
    prof_end(h, st);
    h = prof_begin(4, st);
    msm_merge_affine_kernel<<<(unsigned)((pl.nbuckets + 127) / 128), 128, 0, st>>>(b.offs, s, seg_log, cpw, b.buckets, baff, head, tail, big_count,
                                                                                    big_list, big_cap);
    CQB_LAUNCHED();
    msm_merge_big_affine_kernel<<<big_cap, MERGE_BIG_AFF_THREADS, MERGE_BIG_AFF_THREADS * 128, st>>>(b.offs, s, seg_log, cpw, b.buckets, head, tail,
                                                                                                      big_count, big_list);
    CQB_LAUNCHED();
    prof_end(h, st);
    CQB_CUDA(cudaGetLastError());
    return 0;
}

static int msm_acc_phase(const void* d_bases, size_t n, const MsmShape& s, const PartPlan& pl, const PartBuf& b, cudaStream_t st) {
    // chunking of the bucket-sorted lists: 16..256 entries per chunk thread. One wave is 148 SMs x 4 CTAs x 128 threads = 76k
    // chunks; with fewer than ~10 waves the last, partly filled wave shows (2^22: 213k chunks of 256 = 2.8 waves -> 10.9 ms, 852k
    // chunks of 64 -> 10.3 ms), so the chunks shrink to 64 entries until there are ~1M of them, and further only to keep >= 150k.
    size_t entries = (size_t)n * s.nwin;
    // batched affine: the sorted entries must leave bit 31 free for the first-of-bucket mark
    const bool idx_fits = s.single ? ((size_t)s.nwin * s.table_n < ((size_t)1 << 30)) : true;
    if (idx_fits && g_acc_mode == 2) return msm_acc_phase_affine(d_bases, n, s, pl, b, st);  // measured slower than XYZZ: opt-in only
    int seg_log = 8;
    while (seg_log > 6 && (entries >> seg_log) < 1000000) seg_log--;
    while (seg_log > 4 && (entries >> seg_log) < 150000) seg_log--;
    size_t list_len = s.single ? entries : n;  // entries one set's list can hold for this part
    uint32_t cpw = (uint32_t)((list_len + ((size_t)1 << seg_log) - 1) >> seg_log);  // chunks per set (upper bound)
    size_t nchunks = (size_t)s.nsets * cpw;
    uint32_t big_cap = (uint32_t)(s.nsets * (cpw / MERGE_LONG + 2));
    CQB_TRY(g_chunks->ensure(nchunks * 256 + 16 + (size_t)big_cap * 8));
    uint4* head = g_chunks->as<uint4>();
    uint4* tail = head + nchunks * 8;
    uint32_t* big_count = (uint32_t*)(tail + nchunks * 8);
    uint2* big_list = (uint2*)(big_count + 4);
    CQB_CUDA(cudaMemsetAsync(b.buckets, 0, pl.nbuckets * 128, st));
    CQB_CUDA(cudaMemsetAsync(big_count, 0, 16, st));
    int h = prof_begin(3, st);
    const unsigned acc_grid = (unsigned)((nchunks + 127) / 128);
    msm_accumulate_kernel<false><<<acc_grid, 128, 0, st>>>((const uint4*)d_bases, b.sorted, b.offs, b.hist, b.nonempty, s, seg_log, cpw, b.buckets,
                                                          head, tail);
    CQB_LAUNCHED();
    msm_accumulate_kernel<true><<<acc_grid, 128, 0, st>>>((const uint4*)d_bases, b.sorted, b.offs, b.hist, b.nonempty, s, seg_log, cpw, b.buckets,
                                                         head, tail);
    CQB_LAUNCHED();
    prof_end(h, st);
    h = prof_begin(4, st);
    msm_merge_kernel<<<(unsigned)((pl.nbuckets + 127) / 128), 128, 0, st>>>(b.offs, s, seg_log, cpw, b.buckets, head, tail, big_count, big_list,
                                                                             big_cap);
    CQB_LAUNCHED();
    msm_merge_big_kernel<<<big_cap, 128, 0, st>>>(b.offs, s, seg_log, cpw, b.buckets, head, tail, big_count, big_list);
    CQB_LAUNCHED();
    prof_end(h, st);
    CQB_CUDA(cudaGetLastError());
    return 0;
}

// bucket reduction over the sum of the parts' bucket arrays, window combination, affine normalisation
static int msm_finish(const MsmShape& s, int nparts, void* d_out) {  // nparts is folded to 1 below
    cudaStream_t st = ctx().stream;
    size_t nbuckets = (size_t)s.nsets * s.nb;
    // tpw threads per set, ch buckets each (both powers of two): about 16k threads in total (one warp per SM sub-partition)
    uint32_t tpw = 256u;
    while ((size_t)tpw * 2 * s.nsets <= 16384u) tpw *= 2;
    tpw = std::min<uint32_t>(tpw, s.nb);
    uint32_t ch = s.nb / tpw;                            // >= 1; tpw * ch == nb (powers of two)
    tpw = std::max<uint32_t>(tpw, (uint32_t)RED_CTA);    // pad: threads with t * ch >= nb idle
    uint32_t cps = (tpw + RED_CTA - 1) / RED_CTA;  // CTA partials per set
    uint32_t lvl1 = (cps + 2047) / 2048;           // tree-sum levels over them
    CQB_TRY(g_partials->ensure(((size_t)s.nsets * (cps + lvl1 + 1) + 1) * 128));
    uint4* partials = g_partials->as<uint4>();
    uint4* sums1 = partials + (size_t)s.nsets * cps * 8;
    uint4* wins = sums1 + (size_t)s.nsets * lvl1 * 8;
    int h = prof_begin(5, st);
    if (nparts > 1) {
        msm_fold_parts_kernel<<<(unsigned)((nbuckets + 127) / 128), 128, 0, st>>>(g_buckets->as<uint4>(), nbuckets, nparts);
        CQB_LAUNCHED();
        nparts = 1;
    }
    msm_reduce_kernel<<<(unsigned)(s.nsets * cps), RED_CTA, 0, st>>>(g_buckets->as<uint4>(), s, tpw, ch, nparts, nbuckets * 8, partials);
    CQB_LAUNCHED();
    prof_end(h, st);
    h = prof_begin(6, st);
    if (lvl1 > 1) {
        xyzz_sum_kernel<<<s.nsets * lvl1, 128, 0, st>>>(partials, cps, lvl1, sums1);
        CQB_LAUNCHED();
        xyzz_sum_kernel<<<s.nsets, 128, 0, st>>>(sums1, lvl1, 1, wins);
        CQB_LAUNCHED();
    } else {
        xyzz_sum_kernel<<<s.nsets, 128, 0, st>>>(partials, cps, 1, wins);
        CQB_LAUNCHED();
    }
    prof_end(h, st);
    h = prof_begin(7, st);
    if (s.single && s.nsets > 1) msm_final_batch_kernel<<<s.nsets, 32, 0, st>>>(wins, (uint4*)d_out);
    else msm_final_kernel<<<1, 32, 0, st>>>(wins, s, (uint4*)d_out);
    CQB_LAUNCHED();
    prof_end(h, st);
    CQB_CUDA(cudaGetLastError());
    return 0;
}

// number of parts a device-resident MSM of n points is cut into (sort of part p+1 overlaps the accumulation of part p)
static int g_forced_parts = 0;
void msm_set_parts(int p) { g_forced_parts = p; }
static int auto_parts(size_t n, const MsmShape& s, const uint32_t* d_idx) {
    if (s.nsets > 1 && s.single) return 1;  // batched MSMs: one launch sequence for the whole batch
    if (d_idx) return 1;
    if (g_forced_parts > 0) return std::min(g_forced_parts, MSM_MAX_PARTS);
    // Measured (B200, 2^20..2^24, tools/sweep_msm.py with CQB_PARTS=1/2/4/8): overlapping the sort with the accumulation does
    // not pay for device-resident scalars — the co-resident sort CTAs take register-file space from the accumulate warps and
    // slow them by about the time the sort would have taken alone (2^24: 39.5 / 39.3 / 41.1 / 42.1 ms; re-measured with the
    // geometric part sizes and the bucket-array fold: 38.0 / 38.6 / 38.7 / 40.1 ms). Parts are used only where there is a copy
    // to hide (host-pointer MSM from pinned memory).
    (void)n;
    return 1;
}

// Part boundaries. Device-resident scalars: equal parts. Host-pointer MSM (ready != null): the H2D copy runs ~4x faster
// than the MSM consumes points, so only the FIRST part's copy is exposed — it is made small (1/16 of the points) and the
// rest is split evenly; capi.cu issues its copies with the same function.
void msm_part_bounds(size_t n, int nparts, bool small_first, size_t* bounds) {
    if (nparts < 1) nparts = 1;
    if (nparts > MSM_MAX_PARTS) nparts = MSM_MAX_PARTS;
    bounds[0] = 0;
    if (small_first && nparts >= 3 && n >= ((size_t)1 << 20)) {
        // geometric start: 1/16, then 1/4, then the rest evenly — each part's copy finishes while the previous part computes
        size_t first = n / 16, second = n / 4, rest = n - first - second, per = (rest + (nparts - 2) - 1) / (nparts - 2);
        bounds[1] = first;
        bounds[2] = first + second;
        for (int p = 3; p <= nparts; p++) bounds[p] = std::min(n, first + second + (size_t)(p - 2) * per);
    } else if (small_first && nparts == 2 && n >= ((size_t)1 << 20)) {
        bounds[1] = n / 16;
    } else {
        size_t per = (n + nparts - 1) / nparts;
        for (int p = 1; p <= nparts; p++) bounds[p] = std::min(n, (size_t)p * per);
    }
    bounds[nparts] = n;
}

static int msm_empty_out(void* d_out) {
    static const uint32_t zero_pt[20] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0};
    CQB_CUDA(cudaMemcpyAsync(d_out, zero_pt, sizeof(zero_pt), cudaMemcpyHostToDevice, ctx().stream));
    return 0;
}

// `ready[p]` (optional): event the sort of part p must wait for (the H2D copy of its scalars). `feeder` (optional, instead of
// ready): called on the enqueuing host thread right before part p's sort is queued; it starts the transfer of the part's
// scalars and hands back the event to wait for. Its host-side work (staging pageable memory into pinned buffers) overlaps
// the kernels of the parts already queued: the accumulation of part p-1 is queued before the feeder of part p+1 runs.
static int msm_run_shape(const void* d_bases, const void* d_scalars, const uint32_t* d_idx, size_t n, MsmShape s, int nparts,
                         const cudaEvent_t* ready, MsmFeeder* feeder, void* d_out) {
    cudaStream_t st = ctx().stream;
    g_spans_used = 0;
    if (nparts > MSM_MAX_PARTS) nparts = MSM_MAX_PARTS;
    if (nparts <= 1 && !ready && !feeder) {
        PartPlan pl;
        CQB_TRY(plan_parts(s, 1, &pl));
        PartBuf b = part_buf(s, pl, 0);
        CQB_TRY(msm_sort_phase(d_scalars, d_idx, n, s, pl, b, st));
        CQB_TRY(msm_acc_phase(d_bases, n, s, pl, b, st));
        return msm_finish(s, 1, d_out);
    }
    if (nparts < 1) nparts = 1;
    size_t bounds[MSM_MAX_PARTS + 1];
    msm_part_bounds(n, nparts, ready != nullptr || feeder != nullptr, bounds);
    size_t per = 0;
    for (int p = 0; p < nparts; p++) per = std::max(per, bounds[p + 1] - bounds[p]);
    // per-part list capacity
    s.list_cap = s.single ? per * (size_t)s.nwin : per;
    PartPlan pl;
    CQB_TRY(plan_parts(s, nparts, &pl));
    CQB_TRY(ensure_sort_stream());
    const uint32_t offset0 = s.offset;
    CQB_CUDA(cudaEventRecord(g_ev_start, st));  // the sort stream starts after everything already queued on the main stream
    CQB_CUDA(cudaStreamWaitEvent(g_sort_stream, g_ev_start, 0));
    int used = 0;  // parts that hold points (bucket arrays of the others are never touched and must not be summed)
    size_t prev_lo = 0, prev_cnt = 0;
    auto queue_acc = [&](int slot, size_t lo, size_t cnt) -> int {
        MsmShape sp = s;
        sp.offset = offset0 + (uint32_t)lo;
        PartBuf b = part_buf(s, pl, slot);
        CQB_CUDA(cudaStreamWaitEvent(st, g_ev_sorted[slot], 0));
        return msm_acc_phase(d_bases, cnt, sp, pl, b, st);
    };
    for (int p = 0; p < nparts; p++) {
        size_t lo = bounds[p], cnt = bounds[p + 1] - lo;
        if (feeder) {
            cudaEvent_t ev = nullptr;
            CQB_TRY(feeder->feed(p, lo, cnt, &ev));
            if (ev) CQB_CUDA(cudaStreamWaitEvent(g_sort_stream, ev, 0));
        } else if (ready) {
            CQB_CUDA(cudaStreamWaitEvent(g_sort_stream, ready[p], 0));
        }
        if (cnt == 0) continue;
        MsmShape sp = s;
        sp.offset = offset0 + (uint32_t)lo;
        PartBuf b = part_buf(s, pl, used);
        CQB_TRY(msm_sort_phase((const char*)d_scalars + lo * 32, d_idx ? d_idx + lo : nullptr, cnt, sp, pl, b, g_sort_stream));
        CQB_CUDA(cudaEventRecord(g_ev_sorted[used], g_sort_stream));
        if (used > 0) CQB_TRY(queue_acc(used - 1, prev_lo, prev_cnt));  // the previous part's accumulation, under this part's sort
        prev_lo = lo;
        prev_cnt = cnt;
        used++;
    }
    if (used > 0) CQB_TRY(queue_acc(used - 1, prev_lo, prev_cnt));
    if (used == 0) return msm_empty_out(d_out);
    return msm_finish(s, used, d_out);
}

static int msm_empty(void* d_out) { return msm_empty_out(d_out); }

// windowed layout: sum_i scalars[i] * bases[idx ? idx[i] : offset + i]
int msm_run(const void* d_bases, size_t offset, const void* d_scalars, const uint32_t* d_idx, size_t n, void* d_out, int nparts,
            const cudaEvent_t* ready, MsmFeeder* feeder) {
    if (n == 0) return msm_empty(d_out);  // best_multiexp of empty slices returns the identity
    if (n > ((size_t)1 << 30)) return fail(CQB_E_BAD_SIZE, "MSM of %zu points exceeds the supported 2^30", n);
    MsmShape s = windowed_shape(n);
    s.offset = (uint32_t)offset;
    return msm_run_shape(d_bases, d_scalars, d_idx, n, s, nparts > 0 ? nparts : auto_parts(n, s, d_idx), ready, feeder, d_out);
}

// single-set layout over a precomputed table of `table_n` points per row built with window bits c
int msm_run_precomputed(const void* d_table, size_t table_n, int c, size_t offset, const void* d_scalars, const uint32_t* d_idx, size_t n,
                        void* d_out, int batch, int nparts, const cudaEvent_t* ready, MsmFeeder* feeder) {
    if (n == 0) {
        for (int b = 0; b < batch; b++) CQB_TRY(msm_empty((char*)d_out + (size_t)b * 80));
        return 0;
    }
    MsmShape s;
    s.c = c;
    s.nwin = msm_windows_for(c);
    s.nsets = batch;
    s.single = 1;
    s.nb = 1u << (c - 1);
    s.stride = s.nb + 2;
    s.table_n = (uint32_t)table_n;
    s.offset = (uint32_t)offset;
    s.list_cap = n * (size_t)s.nwin;
    if (s.list_cap >= ((size_t)1 << 32) || (size_t)s.nwin * table_n >= ((size_t)1 << 31))
        return fail(CQB_E_BAD_SIZE, "precomputed MSM: %zu x %d entries exceed the 32-bit index range", n, s.nwin);
    return msm_run_shape(d_table, d_scalars, d_idx, n, s, nparts > 0 ? nparts : auto_parts(n, s, d_idx), ready, feeder, d_out);
}

int g1_batch_normalize_run(const void* d_jacobian, size_t n, void* d_affine) {
    if (n == 0) return 0;
    const size_t threads = (n + NORM_RUN - 1) / NORM_RUN;
    g1_batch_normalize_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx().stream>>>((const uint4*)d_jacobian, n, (uint4*)d_affine);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

int g1_sum_affine_run(const void* d_points, size_t n, void* d_out) {
    g1_sum_affine_kernel<<<1, 128, 0, ctx().stream>>>((const uint4*)d_points, n, (uint4*)d_out);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cqb

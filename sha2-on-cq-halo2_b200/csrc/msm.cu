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
// Random reads of 64 B table records: by default a missing sector pulls its whole 128 B line from DRAM (126.7 B per read measured,
// tools/ubench_gather.cu), the other half of which is a neighbouring point nobody asked for; the L2::64B prefetch-size qualifier
// halves that (63.7 B per read) — the gather-heavy kernels are DRAM-bound otherwise.
__device__ __forceinline__ uint4 ldg_u4_64(const uint4* p) {
    uint4 r;
    asm("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ Fq ldg_fq_64(const uint4* p) {
    uint4 a = ldg_u4_64(p), b = ldg_u4_64(p + 1);
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
    size_t list_cap;  // capacity of one set's sorted list: n (windowed) or n * nwin (single set), plus the padding below
    int pad_log;      // affine-tree accumulation: every bucket's run in the sorted list is padded to a multiple of 2^pad_log entries
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
// Padded layout (s.pad_log > 0, affine-tree accumulation): bucket d's run starts at a multiple of 2^pad_log and its unused
// slots are filled with AFT_PAD here; cursor[] = entry positions (what the scatter advances, cursor[nb+1] = padded total),
// offs[] = the same in units of 2^pad_log entries (what the accumulation over the tree's outputs and the merge kernels read).
constexpr uint32_t AFT_PAD = 0xffffffffu;
__device__ __forceinline__ uint32_t pad_up(uint32_t v, int pad_log) { return (v + ((1u << pad_log) - 1u)) & ~((1u << pad_log) - 1u); }
__global__ void __launch_bounds__(1024) msm_scan_kernel(uint32_t* __restrict__ hist, uint32_t* __restrict__ offs,
                                                        uint32_t* __restrict__ cursor, MsmShape s, uint32_t* __restrict__ nonempty,
                                                        uint32_t* __restrict__ sorted) {
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
        const uint32_t raw = (d <= s.nb) ? h[d] : 0u;
        uint32_t v = pad_up(raw, s.pad_log);
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
        if (d <= s.nb + 1) { o[d] = excl >> s.pad_log; cu[d] = excl; }
        for (uint32_t j = raw; j < v; j++) sorted[(size_t)w * s.list_cap + excl + j] = AFT_PAD;
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
                                                              uint32_t ntiles, int pad_log) {
    __shared__ uint32_t ws[32];
    uint32_t base = 1 + blockIdx.x * SCAN_TILE + threadIdx.x * 8;
    uint32_t v = 0, first = NO_BUCKET;
#pragma unroll
    for (int j = 7; j >= 0; j--) {
        uint32_t d = base + j;
        if (d <= nb) { uint32_t c = hist[d]; v += pad_up(c, pad_log); if (c) first = d; }
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
                                                          uint32_t* __restrict__ nonempty, int pad_log, uint32_t* __restrict__ sorted) {
    __shared__ uint32_t ws[32];
    uint32_t base = 1 + blockIdx.x * SCAN_TILE + threadIdx.x * 8;
    uint32_t vals[8], v = 0, first = NO_BUCKET;
#pragma unroll
    for (int j = 7; j >= 0; j--) {
        uint32_t d = base + j;
        vals[j] = (d <= nb) ? hist[d] : 0u;
        v += pad_up(vals[j], pad_log);
        if (vals[j]) first = d;
    }
    uint32_t excl = block_excl_scan_1024(v, ws, nullptr) + tile_sums[blockIdx.x];
    uint32_t run = min(block_excl_suffix_min_1024(first, ws), tile_sums[ntiles + blockIdx.x]);  // first non-empty id after this thread's 8
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t d = base + j;
        if (d <= nb + 1) { offs[d] = excl >> pad_log; cursor[d] = excl; }
        const uint32_t padded = pad_up(vals[j], pad_log);
        for (uint32_t q = vals[j]; q < padded; q++) sorted[excl + q] = AFT_PAD;
        excl += padded;
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

// (3') PARTITIONED SORT for one large bucket set (single-set layout, c >= 11) — an experiment that did NOT pay (see use_part_sort): kept
// selectable and parity-tested (tests/test_gpu_partitioned_sort.py). The one-thread-per-scalar
// scatter above issues one L2 atomic and one scattered 4 B store per list entry (218 M of each at 2^24: 1.17 ms count + 3.34 ms
// scatter). Here the bucket id (c - 1 bits) is split into 9 coarse + F fine bits:
//   msm_part_kernel      a CTA takes a TILE of 256 x spt scalars, derives their digits and sorts the tile's entries by COARSE bin in
//                        shared memory (histogram, scan, placement: shared-memory atomics only), then writes them out coalesced as
//                        (fine id, entry) pairs, with the tile's 513 bin offsets beside them;
//   msm_part_bin_kernel  work item = (coarse bin, chunk of tiles): reads that bin's segment of every tile of the chunk, counts the fine
//                        ids in shared memory and adds the non-zero counts to the global histogram (COUNT), or — after the same scans
//                        as before — reserves its share of every bucket with one global atomic per bucket and places the entries
//                        (PLACE); all stores of a bin land in that bin's ~1.7 MB window of the list.
// Global atomics drop from 2 per entry to ~2 per (work item, bucket); the order inside a bucket is as arbitrary as before.
constexpr int PS_THREADS = 256;
constexpr int PS_COARSE_LOG = 9;
constexpr int PS_COARSE = 1 << PS_COARSE_LOG;
constexpr int PS_TAB = PS_COARSE + 2;  // uint16 offsets per tile: 513 used, padded to an even count
static inline int part_spt(int nwin) { return nwin <= 16 ? 3 : 2; }  // scalars per thread: 256 x spt x nwin <= 12288 entries per tile
// slot reservation in a shared-memory counter array, warp-aggregated when lanes share a key (hot buckets would serialise otherwise);
// key 0 = no entry. Returns the slot (COUNT_ONLY: nothing).
template <bool WANT_SLOT>
__device__ __forceinline__ uint32_t smem_reserve(uint32_t* counters, uint32_t key) {  // counters[key - 1]
    const uint32_t lane = threadIdx.x & 31u;
    const int mode = warp_key_mode(key);
    uint32_t slot = 0;
    if (mode == 0) return 0;
    if (mode == 2) {
        const uint32_t peers = __match_any_sync(0xffffffffu, key);
        const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
        uint32_t base = 0;
        if (key && lane == leader) base = atomicAdd(counters + (key - 1), (uint32_t)__popc(peers));
        if (WANT_SLOT) slot = __shfl_sync(0xffffffffu, base, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
    } else if (key) {
        slot = atomicAdd(counters + (key - 1), 1u);
    }
    return slot;
}
__global__ void __launch_bounds__(PS_THREADS, 2) msm_part_kernel(const uint4* __restrict__ scalars, const uint32_t* __restrict__ idx, size_t n,
                                                                 MsmShape s, int spt, int fine_bits, uint2* __restrict__ local,
                                                                 uint16_t* __restrict__ tab) {
    extern __shared__ uint4 ps_sm[];
    uint32_t* hist = (uint32_t*)ps_sm;         // [512] counts, then running cursors
    uint32_t* offs = hist + PS_COARSE;         // [513] exclusive offsets
    uint2* ent = (uint2*)(offs + PS_COARSE + 4);  // [cap] the tile's entries in coarse-bin order (1028 words before it: 8 B aligned)
    const uint32_t cap = (uint32_t)(PS_THREADS * spt * s.nwin);
    const size_t tile = blockIdx.x, first = tile * (size_t)(PS_THREADS * spt);
    for (int b = threadIdx.x; b < PS_COARSE; b += PS_THREADS) hist[b] = 0;
    __syncthreads();
    Fr k[3];
    uint32_t pid[3];
    bool act[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const size_t i = first + (size_t)j * PS_THREADS + threadIdx.x;
        act[j] = j < spt && i < n;
        k[j] = act[j] ? load_scalar_canonical(scalars, i) : Fr::zero();
        pid[j] = act[j] ? (idx ? __ldg(idx + i) : (uint32_t)i + s.offset) : 0u;
    }
    const uint32_t fine_mask = (1u << fine_bits) - 1u;
    // pass 1: coarse histogram
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (j >= spt) break;
        uint32_t carry = 0;
        for (int w = 0; w < s.nwin; w++) {
            uint32_t d = window_bits(k[j].l, w, s.c) + carry;
            carry = 0;
            if (d > s.nb) { d = (1u << s.c) - d; carry = 1; }
            const uint32_t key = (act[j] && d) ? ((d - 1u) >> fine_bits) + 1u : 0u;
            smem_reserve<false>(hist, key);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive scan of the 512 counts: 16 per lane + a warp scan
        const int lane = threadIdx.x;
        uint32_t v[16], sum = 0;
#pragma unroll
        for (int q = 0; q < 16; q++) { v[q] = hist[lane * 16 + q]; sum += v[q]; }
        uint32_t x = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        uint32_t run = x - sum;
#pragma unroll
        for (int q = 0; q < 16; q++) { offs[lane * 16 + q] = run; hist[lane * 16 + q] = run; run += v[q]; }
        if (lane == 31) offs[PS_COARSE] = run;
    }
    __syncthreads();
    for (int b = threadIdx.x; b <= PS_COARSE; b += PS_THREADS) tab[tile * PS_TAB + b] = (uint16_t)offs[b];
    // pass 2: placement in shared memory
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (j >= spt) break;
        uint32_t carry = 0;
        for (int w = 0; w < s.nwin; w++) {
            uint32_t d = window_bits(k[j].l, w, s.c) + carry;
            carry = 0;
            uint32_t neg = 0;
            if (d > s.nb) { d = (1u << s.c) - d; carry = 1; neg = 1; }
            const bool has = act[j] && d;
            const uint32_t key = has ? ((d - 1u) >> fine_bits) + 1u : 0u;
            const uint32_t slot = smem_reserve<true>(hist, key);
            if (has) ent[slot] = make_uint2((d - 1u) & fine_mask, ((pid[j] + (uint32_t)w * s.table_n) << 1) | neg);
        }
    }
    __syncthreads();
    const uint32_t total = offs[PS_COARSE];
    uint2* dst = local + tile * (size_t)cap;
    for (uint32_t e = threadIdx.x; e < total; e += PS_THREADS) dst[e] = ent[e];
}
template <bool PLACE>
__global__ void __launch_bounds__(PS_THREADS) msm_part_bin_kernel(const uint2* __restrict__ local, const uint16_t* __restrict__ tab, uint32_t ntiles,
                                                                  uint32_t cap, int fine_bits, uint32_t tiles_per_chunk,
                                                                  uint32_t* __restrict__ counters /* hist (COUNT) or cursor (PLACE), by bucket id */,
                                                                  uint32_t* __restrict__ sorted) {
    extern __shared__ uint4 ps_sm[];
    uint32_t* fh = (uint32_t*)ps_sm;            // [2^F] counts, then running cursors
    uint32_t* base = fh + ((size_t)1 << fine_bits);  // [2^F] PLACE: reserved start of this work item's share of every bucket
    const uint32_t bin = blockIdx.y, nfine = 1u << fine_bits;
    const uint32_t t0 = blockIdx.x * tiles_per_chunk, t1 = min(t0 + tiles_per_chunk, ntiles);
    for (uint32_t f = threadIdx.x; f < nfine; f += PS_THREADS) fh[f] = 0;
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    // every lane walks the segment of ITS OWN tile (32 tiles per warp step): the 32 offset reads and then the 32 entry reads of a step are
    // independent — one tile per warp at a time was a chain of dependent loads per 19-entry segment and ran 8x slower
    for (uint32_t tb = t0 + warp * 32; tb < t1; tb += PS_THREADS) {
        const uint32_t t = tb + lane;
        uint32_t a = 0, e = 0;
        if (t < t1) { a = tab[(size_t)t * PS_TAB + bin]; e = tab[(size_t)t * PS_TAB + bin + 1]; }
        const uint2* seg = local + (size_t)t * cap + a;
        const uint32_t len = e - a, maxlen = __reduce_max_sync(0xffffffffu, len);
        for (uint32_t q = 0; q < maxlen; q++) smem_reserve<false>(fh, q < len ? __ldg(&seg[q].x) + 1u : 0u);
    }
    __syncthreads();
    const uint32_t bucket0 = 1u + (bin << fine_bits);  // bucket id of fine id 0
    if constexpr (!PLACE) {
        for (uint32_t f = threadIdx.x; f < nfine; f += PS_THREADS)
            if (fh[f]) atomicAdd(counters + bucket0 + f, fh[f]);
    } else {
    for (uint32_t f = threadIdx.x; f < nfine; f += PS_THREADS) {
        const uint32_t c = fh[f];
        base[f] = c ? atomicAdd(counters + bucket0 + f, c) : 0u;
        fh[f] = 0;
    }
    __syncthreads();
    for (uint32_t tb = t0 + warp * 32; tb < t1; tb += PS_THREADS) {
        const uint32_t t = tb + lane;
        uint32_t a = 0, e = 0;
        if (t < t1) { a = tab[(size_t)t * PS_TAB + bin]; e = tab[(size_t)t * PS_TAB + bin + 1]; }
        const uint2* seg = local + (size_t)t * cap + a;
        const uint32_t len = e - a, maxlen = __reduce_max_sync(0xffffffffu, len);
        for (uint32_t q = 0; q < maxlen; q++) {
            uint2 x = make_uint2(0, 0);
            if (q < len) x = __ldg(seg + q);
            const uint32_t slot = smem_reserve<true>(fh, q < len ? x.x + 1u : 0u);
            if (q < len) sorted[base[x.x] + slot] = x.y;
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
// DIRECT: the list is the identity map over `bases` (the affine tree's outputs, one point per 2^pad_log entries): no entry load.
template <bool SPARSE, bool DIRECT = false>
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
        uint32_t e = DIRECT ? (pos << 1) : __ldg(lst + pos);
        const uint4* bp = bases + (size_t)(e >> 1) * 4;
#ifdef CQB_XYZZ_GATHER_DEFAULT
        Fq x = ldg_fq(bp), y = ldg_fq(bp + 2);
#else
        Fq x = DIRECT ? ldg_fq(bp) : ldg_fq_64(bp), y = DIRECT ? ldg_fq(bp + 2) : ldg_fq_64(bp + 2);
#endif
        if (x.is_zero() && y.is_zero()) continue;  // identity base contributes nothing (reference curve.rs:857-858)
        if (e & 1u) y = fp_neg<FqP>(y);
        g1_madd(acc, x, y);
    }
    if (is_first) st_xyzz(head + gid * 8, acc);
    else st_xyzz(tail + gid * 8, acc);
}

// (4'') AFFINE-TREE bucket accumulation. The bucket-sorted list is laid out with every bucket's run padded to a multiple of
// S = 2^pad_log entries (AFT_PAD sentinels = identity), so that at every level l = 1..pad_log the pair (2p, 2p+1) of the
// level's input list lies inside ONE bucket and all pair additions of a level are independent: a level is a pass over the whole
// list, 218 M -> 109 M -> ... points at 2^24, and after pad_log levels one affine point per S entries is left; those (1/S of the
// additions) go through the XYZZ chunk kernel (DIRECT) and the usual merge. An affine addition costs 5M + 1S = 788 MAD32 against
// 1,232 for the XYZZ mixed addition because its field inversion is shared by Montgomery's trick, in three launches per level:
//   forward  (aft_forward_tile)   a thread owns T = 32 pairs (interleaved over the CTA for coalescing): denominators x2 - x1, their
//                                 running product parked per pair (32 B), the thread's total, 2 kind bits per pair;
//   invert   (aft_invert_kernel)  Montgomery's trick once more over 32 thread totals per thread and ONE branch-free safegcd inversion
//                                 (fp.cuh) per thread: every lane inverts its own value, 32 x 32 x 32 additions per inversion stream;
//   backward (aft_backward_tile)  peels the inverses off, lambda, x3, y3; 64 B per pair out, coalesced. Level 1 gathers its points
//                                 from the table (four lanes per 64 B record, L2::64B loads), the other levels read the previous
//                                 level's output.
// aft_level_kernel runs the backward role of one slab of tiles and the forward role of the next in one launch. Exceptional pairs
// (reference batch_add, arithmetic/curves/src/derive/curve.rs:4-141, and the mixed addition's branches :866-871): an identity operand
// (sentinel, identity base, an earlier P + (-P)) passes the other one through, P + P joins the batch as a doubling (denominator 2y,
// numerator 3x^2), P + (-P) gives the identity. Identity = (0, 0). Measurements and the variants that lost: profiles/r02_affine_tree.md.
// the two operands of a pair: pointers to their 64 B records (nullptr = sentinel) and sign flags
struct AftPair {
    const uint4 *pa, *pb;
    uint32_t na, nb;
};
template <bool GATHER>
__device__ __forceinline__ uint2 aft_entry(const uint32_t* __restrict__ sorted, size_t p, bool valid) {
    if (GATHER && valid) return __ldg((const uint2*)sorted + p);
    return make_uint2(AFT_PAD, AFT_PAD);
}
template <bool GATHER>
__device__ __forceinline__ AftPair aft_pair(const uint4* __restrict__ pts, uint2 e, size_t p, bool valid) {
    AftPair q;
    if (GATHER) {
        q.pa = e.x == AFT_PAD ? nullptr : pts + (size_t)(e.x >> 1) * 4;
        q.pb = e.y == AFT_PAD ? nullptr : pts + (size_t)(e.y >> 1) * 4;
        q.na = e.x & 1u;
        q.nb = e.y & 1u;
    } else {
        q.pa = valid ? pts + p * 8 : nullptr;
        q.pb = valid ? q.pa + 4 : nullptr;
        q.na = q.nb = 0;
    }
    return q;
}
constexpr int AFT_NT = 128;
#ifndef AFT_FWD_BATCH
#define AFT_FWD_BATCH 4
#endif
__device__ __noinline__ Fq aft_mul(Fq a, Fq b) { return fp_mul<FqP>(a, b); }
#define AFT_P(i) (tile + (size_t)(i) * AFT_NT + threadIdx.x)
// forward: thread g owns the T pairs tile + i * AFT_NT + tid; the running product of their denominators after pair p is parked in
// pf[p] (32 B, coalesced), the thread's total in tot[g], the pair kinds (2 bits each: 0 add, 1 double, 2 pass an operand through,
// 3 P + (-P)) in kinds[g]. Loads run one pair (entries: two pairs) ahead of the arithmetic.
template <bool GATHER, int T>
__device__ __forceinline__ void aft_forward_tile(const uint4* __restrict__ pts, const uint32_t* __restrict__ sorted, size_t npairs, size_t tile_idx,
                                                 uint4* __restrict__ pf, uint4* __restrict__ tot, uint64_t* __restrict__ kinds_out) {
    const size_t tile = tile_idx * (AFT_NT * T);
    if (tile >= npairs) return;
    Fq r = Fq::one();
    uint64_t kinds = 0;
    constexpr int B = AFT_FWD_BATCH;  // pairs whose gathers are in flight together (the kernel is bound by their latency, not by its one multiplication per pair)
    uint2 e[B];
#pragma unroll
    for (int b = 0; b < B; b++) e[b] = aft_entry<GATHER>(sorted, AFT_P(b), AFT_P(b) < npairs);
#pragma unroll 1
    for (int i0 = 0; i0 < T; i0 += B) {
        AftPair q[B];
        Fq xa[B], xb[B];
#pragma unroll
        for (int b = 0; b < B; b++) {
            q[b] = aft_pair<GATHER>(pts, e[b], AFT_P(i0 + b), AFT_P(i0 + b) < npairs);
            xa[b] = q[b].pa ? (GATHER ? ldg_fq_64(q[b].pa) : ldg_fq(q[b].pa)) : Fq::zero();
            xb[b] = q[b].pb ? (GATHER ? ldg_fq_64(q[b].pb) : ldg_fq(q[b].pb)) : Fq::zero();
        }
#pragma unroll
        for (int b = 0; b < B; b++) e[b] = aft_entry<GATHER>(sorted, AFT_P(i0 + B + b), i0 + B + b < T && AFT_P(i0 + B + b) < npairs);
#pragma unroll
        for (int b = 0; b < B; b++) {
            const int i = i0 + b;
            uint32_t kind = 2;
            bool ida = !q[b].pa, idb = !q[b].pb;
            if (q[b].pa && xa[b].is_zero()) ida = ldg_fq(q[b].pa + 2).is_zero();
            if (q[b].pb && xb[b].is_zero()) idb = ldg_fq(q[b].pb + 2).is_zero();
            if (!(ida || idb)) {
                Fq d;
                if (xa[b] == xb[b]) {
                    Fq ya = ldg_fq(q[b].pa + 2), yb = ldg_fq(q[b].pb + 2);
                    if (q[b].na) ya = fp_neg<FqP>(ya);
                    if (q[b].nb) yb = fp_neg<FqP>(yb);
                    if (ya == yb) { kind = 1; d = fp_dbl<FqP>(ya); }
                    else kind = 3;
                } else {
                    kind = 0;
                    d = fp_sub<FqP>(xb[b], xa[b]);
                }
                if (kind < 2) r = aft_mul(r, d);
            }
            kinds |= (uint64_t)kind << (2 * i);
            if (AFT_P(i) < npairs) st_fq(pf + AFT_P(i) * 2, r);
        }
    }
    const size_t g = tile_idx * AFT_NT + threadIdx.x;
    st_fq(tot + g * 2, r);
    kinds_out[g] = kinds;
}
// inversion of the thread totals [first, first + count) in place (count is clipped to what the forward tiles wrote): thread j owns the
// totals first + j, first + j + stride, ... (coalesced), Montgomery's trick over them (running products in tp), ONE branch-free safegcd
// inversion per thread — every lane inverts its own value
__global__ void __launch_bounds__(128) aft_invert_kernel(uint4* __restrict__ tot, uint4* __restrict__ tp, const uint32_t* __restrict__ total_entries,
                                                         int level, int pairs_per_cta, size_t first, size_t count, int group) {
    const size_t npairs = (size_t)(*total_entries) >> level;
    const size_t written = (npairs + pairs_per_cta - 1) / pairs_per_cta * AFT_NT;  // totals the forward tiles wrote
    if (first >= written) return;
    const size_t m = min(count, written - first);
    const size_t stride = (m + group - 1) / group;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= stride) return;
    tot += first * 2;
    tp += first * 2;
    Fq acc = ld_fq(tot + j * 2);
    st_fq(tp + j * 2, acc);
    int cnt = 1;
#pragma unroll 1
    for (size_t k = j + stride; k < m; k += stride, cnt++) {
        acc = fp_mul<FqP>(acc, ld_fq(tot + k * 2));
        st_fq(tp + k * 2, acc);
    }
    Fq inv = fp_inv_safegcd<FqP>(acc);
#pragma unroll 1
    for (int c = cnt - 1; c >= 1; c--) {
        const size_t k = j + (size_t)c * stride;
        const Fq v = ld_fq(tot + k * 2);
        st_fq(tot + k * 2, fp_mul<FqP>(inv, ld_fq(tp + (k - stride) * 2)));
        inv = fp_mul<FqP>(inv, v);
    }
    st_fq(tot + j * 2, inv);
}
// Raw 64 B records of a pair's two operands as this lane received them from the loads. GATHER: the four lanes 4j .. 4j+3 read the
// record of one owner lane TOGETHER (16 B each, one coalesced 64 B request per record; a thread reading its own record issues four
// 16 B requests that reach the L2 as two sector requests — random L2 misses are limited to ~47 G requests/s on B200 whatever their
// size, tools/ubench_gather.cu: 2.80 ms -> 1.40 ms for 2^26 random 64 B reads), round k serves the owners 8k .. 8k+7; aft_exchange
// hands the quarters to their owners through shared memory. Sequential levels read their own records (already coalesced per line).
struct AftRaw { uint4 a[4], b[4]; };
template <bool GATHER>
__device__ __forceinline__ void aft_load_raw(AftRaw& w, const AftPair& q) {
    if (GATHER) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int owner = (lane >> 2) + 8 * k;
            const uint4* pa = (const uint4*)__shfl_sync(0xffffffffu, (unsigned long long)q.pa, owner);
            const uint4* pb = (const uint4*)__shfl_sync(0xffffffffu, (unsigned long long)q.pb, owner);
            w.a[k] = pa ? ldg_u4_64(pa + (lane & 3)) : make_uint4(0, 0, 0, 0);
            w.b[k] = pb ? ldg_u4_64(pb + (lane & 3)) : make_uint4(0, 0, 0, 0);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            w.a[k] = q.pa ? __ldg(q.pa + k) : make_uint4(0, 0, 0, 0);
            w.b[k] = q.pb ? __ldg(q.pb + k) : make_uint4(0, 0, 0, 0);
        }
    }
}
__device__ __forceinline__ Fq fq_of(const uint4& lo, const uint4& hi) {
    Fq r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
    r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}
// xch: this warp's 32 x 64 B exchange buffer (quarter-major: [quarter][owner lane], conflict-free 16 B accesses)
template <bool GATHER>
__device__ __forceinline__ void aft_exchange(const AftRaw& w, uint4* xch, Fq& xa, Fq& ya, Fq& xb, Fq& yb) {
    if (GATHER) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int k = 0; k < 4; k++) xch[(lane & 3) * 32 + (lane >> 2) + 8 * k] = w.a[k];
        __syncwarp();
        xa = fq_of(xch[lane], xch[32 + lane]);
        ya = fq_of(xch[64 + lane], xch[96 + lane]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; k++) xch[(lane & 3) * 32 + (lane >> 2) + 8 * k] = w.b[k];
        __syncwarp();
        xb = fq_of(xch[lane], xch[32 + lane]);
        yb = fq_of(xch[64 + lane], xch[96 + lane]);
        __syncwarp();
    } else {
        xa = fq_of(w.a[0], w.a[1]); ya = fq_of(w.a[2], w.a[3]);
        xb = fq_of(w.b[0], w.b[1]); yb = fq_of(w.b[2], w.b[3]);
    }
}
// backward: peel the inverses off, finish the additions (points of pair i-1 and entries of pair i-2 in flight), 64 B per pair out.
// All 32 lanes of a warp run every iteration (the cooperative loads need them); lanes past the end of the list carry null pairs.
template <bool GATHER, int T>
__device__ __forceinline__ void aft_backward_tile(const uint4* __restrict__ pts, const uint32_t* __restrict__ sorted, size_t npairs, size_t tile_idx,
                                                  const uint4* __restrict__ pf, const uint4* __restrict__ tot,
                                                  const uint64_t* __restrict__ kinds_in, uint4* __restrict__ out, uint4* xch_all) {
    const size_t tile = tile_idx * (AFT_NT * T);
    if (tile >= npairs) return;
    uint4* xch = xch_all + (threadIdx.x >> 5) * 128;
    const size_t g = tile_idx * AFT_NT + threadIdx.x;
    Fq inv = ld_fq(tot + g * 2);
    const uint64_t kinds = kinds_in[g];
    AftPair nq = aft_pair<GATHER>(pts, aft_entry<GATHER>(sorted, AFT_P(T - 1), AFT_P(T - 1) < npairs), AFT_P(T - 1), AFT_P(T - 1) < npairs);
    AftRaw raw;
    aft_load_raw<GATHER>(raw, nq);
    uint2 ne = aft_entry<GATHER>(sorted, AFT_P(T > 1 ? T - 2 : 0), T > 1 && AFT_P(T - 2) < npairs);
    Fq npre = (T > 1 && AFT_P(T - 1) < npairs) ? ld_fq(pf + AFT_P(T - 2) * 2) : Fq::one();
#pragma unroll 1  // unrolled by 2: 144-150 registers, 3 CTAs per SM, 30.0 ms against 27.7 ms for the phase at 2^24
    for (int i = T - 1; i >= 0; i--) {
        const AftPair q = nq;
        Fq xa, ya, xb, yb;
        aft_exchange<GATHER>(raw, xch, xa, ya, xb, yb);
        const Fq pre = npre;
        if (i > 0) {
            nq = aft_pair<GATHER>(pts, ne, AFT_P(i - 1), AFT_P(i - 1) < npairs);
            aft_load_raw<GATHER>(raw, nq);
            ne = aft_entry<GATHER>(sorted, AFT_P(i > 1 ? i - 2 : 0), i > 1 && AFT_P(i - 2) < npairs);
            npre = (i > 1 && AFT_P(i - 1) < npairs) ? ld_fq(pf + AFT_P(i - 2) * 2) : Fq::one();
        }
        const size_t p = AFT_P(i);
        if (p < npairs) {
            const uint32_t kind = (uint32_t)(kinds >> (2 * i)) & 3u;
            if (GATHER) {
                if (q.na) ya = fp_neg<FqP>(ya);  // fp_neg(0) = 0: the identity stays (0, 0)
                if (q.nb) yb = fp_neg<FqP>(yb);
            }
            Fq ox, oy;
            if (kind >= 2) {
                if (kind == 3) { ox = Fq::zero(); oy = Fq::zero(); }
                else if (xa.is_zero() && ya.is_zero()) { ox = xb; oy = yb; }
                else { ox = xa; oy = ya; }
            } else {
                Fq d, num;
                if (kind == 0) { d = fp_sub<FqP>(xb, xa); num = fp_sub<FqP>(yb, ya); }
                else {
                    d = fp_dbl<FqP>(ya);
                    const Fq xx = fp_sqr<FqP>(xa);
                    num = fp_add<FqP>(fp_dbl<FqP>(xx), xx);
                }
                Fq dinv = inv;
                if (i > 0) {
                    dinv = fp_mul<FqP>(inv, pre);
                    inv = fp_mul<FqP>(inv, d);
                }
                const Fq lam = fp_mul<FqP>(num, dinv);
                ox = fp_sub<FqP>(fp_sub<FqP>(fp_sqr<FqP>(lam), xa), xb);
                oy = fp_sub<FqP>(fp_mul<FqP>(lam, fp_sub<FqP>(xa, ox)), ya);
            }
            st_fq(out + p * 4, ox);
            st_fq(out + p * 4 + 2, oy);
        }
    }
}
// One launch = the backward pass over the tiles [bwd0, bwd0 + nbwd) and the forward pass over [fwd0, fwd0 + nfwd) of a level, CTAs of the
// two roles interleaved (even / odd blockIdx). The forward pass is bound by its gathers (one multiplication per pair), the backward pass
// by the multiplier pipe; interleaving a later slab's forward pass with an earlier slab's backward pass buys 0.6 ms of the 28 at 2^24 with
// two slabs and nothing beyond (both roles are limited by the same 16 warps per SM; profiles/r02_affine_tree.md).
// 4 CTAs per SM (<= 128 registers, no spills): 27.7 ms for the phase at 2^24; left to the compiler 28.2 ms (124 registers + a spill), forced to 5
// CTAs (96 registers, 300-500 B of spills) 30.5 ms.
template <bool GATHER, int T>
__global__ void __launch_bounds__(AFT_NT, 4) aft_level_kernel(const uint4* __restrict__ pts, const uint32_t* __restrict__ sorted,
                                                          const uint32_t* __restrict__ total_entries, int level, unsigned fwd0, unsigned nfwd,
                                                          unsigned bwd0, unsigned nbwd, uint4* __restrict__ pf, uint4* __restrict__ tot,
                                                          uint64_t* __restrict__ kinds, uint4* __restrict__ out) {
    __shared__ uint4 xch[GATHER ? (AFT_NT / 32) * 128 : 1];
    const size_t npairs = (size_t)(*total_entries) >> level;
    const unsigned idx = blockIdx.x >> 1;
    if ((blockIdx.x & 1u) == 0) {
        if (idx < nbwd) aft_backward_tile<GATHER, T>(pts, sorted, npairs, (size_t)bwd0 + idx, pf, tot, kinds, out, xch);
    } else {
        if (idx < nfwd) aft_forward_tile<GATHER, T>(pts, sorted, npairs, (size_t)fwd0 + idx, pf, tot, kinds);
    }
}
#undef AFT_P

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
static PerDevice<Scratch> g_hist, g_sorted, g_buckets, g_partials, g_chunks, g_pre_tmp, g_aff, g_part;

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

// the grow-only working buffers of the current device (a following MSM reallocates what it needs)
void msm_release_scratch() {
    cudaStreamSynchronize(ctx().stream);
    if (g_sort_stream) cudaStreamSynchronize(g_sort_stream);
    g_hist->release();
    g_sorted->release();
    g_buckets->release();
    g_partials->release();
    g_chunks->release();
    g_pre_tmp->release();
    g_aff->release();
    g_part->release();
}
// working set of one MSM of n points over a single-set table with nwin windows, beyond the scalars: sorted list, tree scratch (parts
// are capped at TREE_PART_ENTRIES entries), histograms / bucket arrays / chunk partials
constexpr size_t TREE_PART_ENTRIES = (size_t)232 << 20;  // 2^24 points x 13 windows + padding
size_t msm_working_set_bytes(size_t n, int nwin) {
    const size_t entries = n * (size_t)nwin, part = std::min(entries, TREE_PART_ENTRIES);
    return entries * 4 + part * 72 + ((size_t)2 << 30);
}

void msm_release_all() {
    g_hist->release();
    g_sorted->release();
    g_buckets->release();
    g_partials->release();
    g_chunks->release();
    g_pre_tmp->release();
    g_aff->release();
    g_part->release();
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
    s.pad_log = 0;
    s.list_cap = n;
    return s;
}

// An MSM runs as `nparts` PARTS over contiguous point ranges (1 part unless it is large). Each part has its own histogram
// / offsets / sorted list / bucket array; its SORT phase (count, scan, scatter) may run on the sort stream while the
// ACCUMULATE phase (accumulate, merge) of the previous part runs on the main stream. msm_finish adds the parts' bucket
// arrays while it reduces them. For a host-pointer MSM the sort of part p additionally waits for the H2D copy of part p.
// capacity of one set's sorted list: the entries plus, in the padded layout, up to 2^pad_log - 1 sentinels per bucket
static size_t padded_cap(size_t entries, const MsmShape& s) {
    if (s.pad_log <= 0) return entries;
    const size_t m = ((size_t)1 << s.pad_log) - 1;
    return (entries + (size_t)s.nb * m + m) & ~m;
}
// accumulation variant: 0 = automatic, 1 = XYZZ mixed additions, 2 = batched affine streams (measured slower), 3 = affine tree
static int g_acc_mode = 0;
void msm_set_accumulator(int mode) { g_acc_mode = mode; }
static int g_tree_levels = 4;  // experiments: levels of the affine tree (entries per padded run = 2^levels)
void msm_set_tree_levels(int l) { g_tree_levels = l < 1 ? 1 : (l > 6 ? 6 : l); }
// levels of the affine tree for an MSM (or a part of one) of n points over shape s; 0 = plain XYZZ accumulation. Measured on B200
// (profiles/r02_affine_tree.md): the tree pays from ~40 entries per bucket on (2^21 points at c = 20: 5.53 vs 5.73 ms; 2^24: 34.0 vs
// 37.9 ms), deeper trees as the buckets get longer; below that the padding and the extra launches cost more than the cheaper additions.
static int tree_pad_log(size_t n, const MsmShape& s) {
    if (g_acc_mode == 1 || g_acc_mode == 2) return 0;
    if (!s.single || s.nsets != 1) return 0;
    const size_t per_bucket = n * (size_t)s.nwin / s.nb;
    if (g_acc_mode == 3) return per_bucket >= ((size_t)4 << g_tree_levels) ? g_tree_levels : 0;
    // ... and from ~24 M list entries on: below that the levels are too short to fill the machine (2^20 points at c = 17, 15.7 M entries:
    // 3.48 ms with the tree against 3.28 ms without)
    if (n * (size_t)s.nwin < ((size_t)24 << 20)) return 0;
    return per_bucket >= 256 ? 4 : per_bucket >= 128 ? 3 : per_bucket >= 40 ? 2 : 0;
}

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

// partitioned sort (kernel comment above): tile pass, per-bin count, the usual scans, per-bin placement
static int g_sort_mode = 0;  // experiments / tests: 0 = automatic, 1 = one-thread-per-scalar scatter, 2 = partitioned whenever the shape allows
void msm_set_sort_mode(int m) { g_sort_mode = m; }
static bool use_part_sort(size_t n, const MsmShape& s) {
    if (g_sort_mode == 1 || !s.single || s.nsets != 1 || s.c < PS_COARSE_LOG + 2 || s.c > 22) return false;
    (void)n;
    return g_sort_mode == 2;  // measured SLOWER than the scatter (2^24: 2.1 + 1.0 + 3.1 ms against 1.17 + 3.34 ms, profiles/r02_partitioned_sort.md): opt-in only
}
static size_t part_scratch_bytes(size_t n, const MsmShape& s) {
    const int spt = part_spt(s.nwin);
    const size_t ntiles = (n + (size_t)PS_THREADS * spt - 1) / ((size_t)PS_THREADS * spt);
    return ntiles * ((size_t)PS_THREADS * spt * s.nwin * 8 + PS_TAB * 2) + 64;
}
static int msm_sort_phase_partitioned(const void* d_scalars, const uint32_t* d_idx, size_t n, const MsmShape& s, const PartPlan& pl, const PartBuf& b,
                                      cudaStream_t st) {
    const int spt = part_spt(s.nwin), fine_bits = s.c - 1 - PS_COARSE_LOG;
    const uint32_t cap = (uint32_t)(PS_THREADS * spt * s.nwin);
    const size_t ntiles = (n + (size_t)PS_THREADS * spt - 1) / ((size_t)PS_THREADS * spt);
    CQB_TRY(g_part->ensure(part_scratch_bytes(n, s)));
    uint2* local = g_part->as<uint2>();
    uint16_t* tab = (uint16_t*)(local + ntiles * cap);
    const size_t smem1 = 1028 * 4 + (size_t)cap * 8, smem2 = ((size_t)2 << fine_bits) * 4;
    static bool attr[MAX_DEVICES] = {};
    if (!attr[cur_slot()]) {
        CQB_CUDA(cudaFuncSetAttribute(msm_part_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1028 * 4 + 12288 * 8));
        CQB_CUDA(cudaFuncSetAttribute(msm_part_bin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CQB_CUDA(cudaFuncSetAttribute(msm_part_bin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr[cur_slot()] = true;
    }
    // work items per coarse bin: chunks of tiles holding ~8 k entries of the bin when the digits are uniform
    const uint32_t per_tile = std::max<uint32_t>(1u, cap / PS_COARSE);
    uint32_t tiles_per_chunk = std::max<uint32_t>(32u, 8192u / per_tile);
    uint32_t chunks = (uint32_t)((ntiles + tiles_per_chunk - 1) / tiles_per_chunk);
    if (chunks > 128u) { chunks = 128u; tiles_per_chunk = (uint32_t)((ntiles + chunks - 1) / chunks); chunks = (uint32_t)((ntiles + tiles_per_chunk - 1) / tiles_per_chunk); }
    const dim3 grid2(chunks, PS_COARSE);
    int h = prof_begin(0, st);
    msm_part_kernel<<<(unsigned)ntiles, PS_THREADS, smem1, st>>>((const uint4*)d_scalars, d_idx, n, s, spt, fine_bits, local, tab);
    CQB_LAUNCHED();
    msm_part_bin_kernel<false><<<grid2, PS_THREADS, smem2, st>>>(local, tab, (uint32_t)ntiles, cap, fine_bits, tiles_per_chunk, b.hist, nullptr);
    CQB_LAUNCHED();
    prof_end(h, st);
    h = prof_begin(1, st);
    if (s.nb > 32768) {
        scan_tile_sums_kernel<<<pl.ntiles, 1024, 0, st>>>(b.hist, s.nb, b.tile_sums, pl.ntiles, s.pad_log);
        CQB_LAUNCHED();
        scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(b.tile_sums, pl.ntiles);
        CQB_LAUNCHED();
        scan_apply_kernel<<<pl.ntiles, 1024, 0, st>>>(b.hist, s.nb, b.tile_sums, pl.ntiles, b.offs, b.cursor, b.nonempty, s.pad_log, b.sorted);
        CQB_LAUNCHED();
    } else {
        msm_scan_kernel<<<1, 1024, 0, st>>>(b.hist, b.offs, b.cursor, s, b.nonempty, b.sorted);
        CQB_LAUNCHED();
    }
    prof_end(h, st);
    h = prof_begin(2, st);
    msm_part_bin_kernel<true><<<grid2, PS_THREADS, smem2, st>>>(local, tab, (uint32_t)ntiles, cap, fine_bits, tiles_per_chunk, b.cursor, b.sorted);
    CQB_LAUNCHED();
    prof_end(h, st);
    CQB_CUDA(cudaGetLastError());
    return 0;
}

// d_bases: windowed layout -> element 0 of the registered set (pid = offset + i); single-set layout -> the table base.
static int msm_sort_phase(const void* d_scalars, const uint32_t* d_idx, size_t n, const MsmShape& s, const PartPlan& pl, const PartBuf& b,
                          cudaStream_t st) {
    CQB_CUDA(cudaMemsetAsync(b.hist, 0, pl.hist_words * 4, st));
    CQB_CUDA(cudaMemsetAsync(b.nonempty, 0, 4, st));
    if (use_part_sort(n, s)) return msm_sort_phase_partitioned(d_scalars, d_idx, n, s, pl, b, st);
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
            scan_tile_sums_kernel<<<pl.ntiles, 1024, 0, st>>>(b.hist + o, s.nb, b.tile_sums, pl.ntiles, s.pad_log);
            CQB_LAUNCHED();
            scan_tile_offsets_kernel<<<1, 1024, 0, st>>>(b.tile_sums, pl.ntiles);
            CQB_LAUNCHED();
            scan_apply_kernel<<<pl.ntiles, 1024, 0, st>>>(b.hist + o, s.nb, b.tile_sums, pl.ntiles, b.offs + o, b.cursor + o, b.nonempty, s.pad_log,
                                                          b.sorted + (size_t)k * s.list_cap);
            CQB_LAUNCHED();
        }
    } else {
        msm_scan_kernel<<<s.nsets, 1024, 0, st>>>(b.hist, b.offs, b.cursor, s, b.nonempty, b.sorted);
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

// affine-tree accumulation (kernel comment above): pad_log levels of pair additions, then the XYZZ chunk kernel over the
// remaining 1 / 2^pad_log of the list; the XYZZ bucket array comes out as msm_acc_phase leaves it
// scratch of the affine tree: level outputs (cap/2 + cap/4 points), parked products (32 B per level-1 pair), thread totals and
// their running products (the forward kernel's threads own >= 8 pairs each), kinds
static size_t tree_scratch_bytes(const MsmShape& s) {
    const size_t cap = s.list_cap, thr = cap / 2 / 8 + 2 * AFT_NT;
    return (cap / 2 + cap / 4) * 64 + (cap / 2) * 32 + thr * (32 + 32 + 8) + 64;
}
static int g_last_tree_levels = 0;  // tree depth the most recent MSM ran with (its largest part); 0 = XYZZ
int msm_last_tree_levels() { return g_last_tree_levels; }
static int g_tree_slabs = 0;  // experiments: slabs per level (0 = automatic)
void msm_set_tree_slabs(int slabs) { g_tree_slabs = slabs; }
template <int T>
static int tree_level_launch(const uint4* in, const uint32_t* sorted, const uint32_t* total, int level, size_t pairs_ub, uint4* pf, uint4* tot,
                             uint4* tp, uint64_t* kinds, uint4* out, cudaStream_t st) {
    const size_t per_cta = (size_t)AFT_NT * T;
    const unsigned tiles = (unsigned)((pairs_ub + per_cta - 1) / per_cta);
    // slabs: the forward pass of slab s+1 shares a launch with the backward pass of slab s (their inversion in between)
    unsigned slabs = g_tree_slabs > 0 ? (unsigned)g_tree_slabs : ((size_t)tiles * T >= 256000u ? 2u : 1u);  // measured: 2 slabs 28.1 ms, 1: 28.7, 4: 28.7, 8: 29.3 (2^24)
    const unsigned per_slab = (tiles + slabs - 1) / slabs;
    slabs = (tiles + per_slab - 1) / per_slab;
    const int group = 32;
    auto launch = [&](unsigned f0, unsigned nf, unsigned b0, unsigned nb) {
        const unsigned grid = 2 * std::max(nf, nb);
        if (sorted) aft_level_kernel<true, T><<<grid, AFT_NT, 0, st>>>(in, sorted, total, level, f0, nf, b0, nb, pf, tot, kinds, out);
        else aft_level_kernel<false, T><<<grid, AFT_NT, 0, st>>>(in, nullptr, total, level, f0, nf, b0, nb, pf, tot, kinds, out);
        CQB_LAUNCHED();
    };
    auto count_of = [&](unsigned sidx) { return std::min(per_slab, tiles - sidx * per_slab); };
    for (unsigned sidx = 0; sidx <= slabs; sidx++) {
        const unsigned nf = sidx < slabs ? count_of(sidx) : 0, nb = sidx > 0 ? count_of(sidx - 1) : 0;
        launch(sidx * per_slab, nf, sidx ? (sidx - 1) * per_slab : 0, nb);
        if (nf) {
            const size_t cnt = (size_t)nf * AFT_NT, inv_threads = (cnt + group - 1) / group;
            aft_invert_kernel<<<(unsigned)((inv_threads + 127) / 128), 128, 0, st>>>(tot, tp, total, level, (int)per_cta, (size_t)sidx * per_slab * AFT_NT,
                                                                                     cnt, group);
            CQB_LAUNCHED();
        }
    }
    return 0;
}
static int msm_acc_phase_tree(const void* d_bases, size_t n, const MsmShape& s, const PartPlan& pl, const PartBuf& b, cudaStream_t st) {
    (void)n;
    const size_t cap = s.list_cap;  // padded capacity, a multiple of 2^pad_log
    CQB_TRY(g_aff->ensure(tree_scratch_bytes(s)));
    uint4* bufA = g_aff->as<uint4>();        // levels 1, 3, ...: cap / 2 points
    uint4* bufB = bufA + (cap / 2) * 4;      // levels 2, 4, ...: cap / 4 points
    uint4* pf = bufB + (cap / 4) * 4;        // parked running products: 32 B per pair
    const size_t thr = cap / 2 / 8 + 2 * AFT_NT;
    uint4* tot = pf + (cap / 2) * 2;
    uint4* tp = tot + thr * 2;
    uint64_t* kinds = (uint64_t*)(tp + thr * 2);
    const uint32_t* total = b.cursor + (size_t)s.nb + 1;  // padded entry total of this part (never advanced by the scatter)
    int h = prof_begin(3, st);
    const uint4* in = (const uint4*)d_bases;
    for (int level = 1; level <= s.pad_log; level++) {
        uint4* out = (level & 1) ? bufA : bufB;
        const size_t pairs_ub = cap >> level;
        const uint32_t* srt = level == 1 ? b.sorted : nullptr;
        CQB_TRY(tree_level_launch<32>(in, srt, total, level, pairs_ub, pf, tot, tp, kinds, out, st));  // 32 pairs per thread (16: +0.2 ms, 8: +0.7 ms at 2^24)
        in = out;
    }
    // the remaining list: cap >> pad_log points, bucket d owns [offs[d], offs[d+1])
    const size_t entries = cap >> s.pad_log;
    int seg_log = 6;
    while (seg_log > 3 && (entries >> seg_log) < 300000) seg_log--;
    uint32_t cpw = (uint32_t)((entries + ((size_t)1 << seg_log) - 1) >> seg_log);
    size_t nchunks = cpw;
    uint32_t big_cap = (uint32_t)(cpw / MERGE_LONG + 2);
    CQB_TRY(g_chunks->ensure(nchunks * 256 + 16 + (size_t)big_cap * 8));
    uint4* head = g_chunks->as<uint4>();
    uint4* tail = head + nchunks * 8;
    uint32_t* big_count = (uint32_t*)(tail + nchunks * 8);
    uint2* big_list = (uint2*)(big_count + 4);
    CQB_CUDA(cudaMemsetAsync(b.buckets, 0, pl.nbuckets * 128, st));
    CQB_CUDA(cudaMemsetAsync(big_count, 0, 16, st));
    const unsigned acc_grid = (unsigned)((nchunks + 127) / 128);
    msm_accumulate_kernel<false, true><<<acc_grid, 128, 0, st>>>(in, nullptr, b.offs, b.hist, b.nonempty, s, seg_log, cpw, b.buckets, head, tail);
    CQB_LAUNCHED();
    msm_accumulate_kernel<true, true><<<acc_grid, 128, 0, st>>>(in, nullptr, b.offs, b.hist, b.nonempty, s, seg_log, cpw, b.buckets, head, tail);
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

static int msm_acc_phase(const void* d_bases, size_t n, const MsmShape& s, const PartPlan& pl, const PartBuf& b, cudaStream_t st) {
    // chunking of the bucket-sorted lists: 16..256 entries per chunk thread. One wave is 148 SMs x 4 CTAs x 128 threads = 76k
    // chunks; with fewer than ~10 waves the last, partly filled wave shows (2^22: 213k chunks of 256 = 2.8 waves -> 10.9 ms, 852k
    // chunks of 64 -> 10.3 ms), so the chunks shrink to 64 entries until there are ~1M of them, and further only to keep >= 150k.
    if (s.pad_log > 0) return msm_acc_phase_tree(d_bases, n, s, pl, b, st);
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
    // the affine tree's scratch is 64 B per list entry: very large MSMs are cut into parts of at most 2^24 x 13 entries so that it stays
    // ~15 GB (2^26: 4 parts; the parts' sorts run under the previous part's accumulation)
    if (tree_pad_log(n, s) > 0) {
        const size_t entries = n * (size_t)s.nwin;
        return (int)std::min<size_t>(MSM_MAX_PARTS, (entries + TREE_PART_ENTRIES - 1) / TREE_PART_ENTRIES);
    }
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
        if (s.pad_log > 0 && g_aff->ensure(tree_scratch_bytes(s)) != 0) {  // no room for the tree's scratch: XYZZ accumulation
            s.pad_log = 0;
            s.list_cap = s.single ? n * (size_t)s.nwin : n;
        }
        g_last_tree_levels = s.pad_log;
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
    // per-part list capacity; each part picks its own tree depth from its size (the padded capacity is that of the largest part)
    s.pad_log = tree_pad_log(per, s);
    s.list_cap = padded_cap(s.single ? per * (size_t)s.nwin : per, s);
    if (s.pad_log > 0 && g_aff->ensure(tree_scratch_bytes(s)) != 0) {
        s.pad_log = 0;
        s.list_cap = s.single ? per * (size_t)s.nwin : per;
    }
    const bool tree_ok = s.pad_log > 0;
    g_last_tree_levels = s.pad_log;
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
        sp.pad_log = tree_ok ? tree_pad_log(cnt, s) : 0;
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
        sp.pad_log = tree_ok ? tree_pad_log(cnt, s) : 0;
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
    s.pad_log = tree_pad_log(n, s);
    s.list_cap = padded_cap(n * (size_t)s.nwin, s);
    if (s.list_cap >= ((size_t)1 << 32) || (size_t)s.nwin * table_n >= ((size_t)1 << 31) - 1)
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

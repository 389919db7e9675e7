// ntt.cu — Fr radix-2 NTT for bn256, replacing best_fft::<Fr> (reference halo2_proofs/src/arithmetic.rs:171-274) and
// the EvaluationDomain wrappers around it (poly/domain.rs:252-266 coeff_to_extended, :293-315 extended_to_coeff,
// :319-338 divide_by_vanishing_poly, :366-374 ifft).
//
// Same mathematical object as the reference: out[k] = sum_j a[j] omega^(jk), natural order in and out, every value
// canonical in [0, r) — so the output limbs are identical to the reference's regardless of how the butterflies are
// scheduled. The schedule here is B200-shaped, not the reference's recursion:
//   * decimation in time over a device-resident twiddle table W[i] = omega^i, i < n/2 (the reference's `twiddles`
//     vector, arithmetic.rs:194-200, built in parallel instead of by a serial scan, and cached per (omega, log_n));
//   * ceil(log_n / 8) passes over HBM; each pass runs up to 8 butterfly stages on a 1024-element (32 KB) tile held in
//     shared memory as two uint4 planes (conflict-free 128-bit LDS/STS); 256 threads, each carrying 4 elements through
//     2 stages in registers (radix-4 steps: half the shared-memory round trips and barriers, 2 independent modmuls);
//   * the bit-reversal permutation (arithmetic.rs:186-191) is fused into the first pass's gather — tiles are chosen so
//     that the gather reads 128 B contiguous runs; later passes read/write (rows x 2^q contiguous elements) tiles;
//   * the element-wise scalings that surround best_fft in the reference are fused into the first pass's load (coset
//     powers, zero padding, division by the vanishing polynomial) and the last pass's store (1/n, coset powers out).
// Roofline: integer-pipe bound ((n/2) log n modmuls at ~136 IMAD.WIDE each); HBM traffic is 64 B/element/pass.
#include <map>
#include <mutex>
#include <vector>

#include "internal.h"

namespace cqb {

#ifndef NTT_MIN_CTAS
#define NTT_MIN_CTAS 3
#endif
constexpr int NTT_TILE_LOG = 10;   // elements per CTA tile (log2)
constexpr int NTT_MAX_R = 8;       // max butterfly stages per pass
constexpr int NTT_PRE_MAX = NTT_PRE_MAX_PUB;

struct NttPassArgs {
    const uint4* src;
    uint4* dst;
    const uint4* tw;
    int L, s0, r, q;
    int first;
    unsigned long long n_in;
    int pre_mode;   // 0 none | 1: x *= pre[j % 3] for j % 3 != 0 | 2: x *= pre[j & (pre_len - 1)]
    int pre_len;
    int post_mode;  // 0 none | 1: x *= post[0] | 2: x *= post[i % 3]
    Fr pre[NTT_PRE_MAX];
    Fr post[3];
    // Address maps of the distributed four-step NTT (batched transforms only). Element idx of member b lives at
    //   (idx >> s) * A + b * B + (idx & (2^s - 1))
    // in_map : the FIRST pass gathers straight out of an all-to-all receive buffer [source rank][member][segment]
    //          (s = log2 segment, A = batch * segment, B = segment) — no interleaving copy;
    // out_map: the LAST pass stores transposed, [idx][member] (s = 0, A = batch, B = 1), i.e. already in the
    //          [destination rank][...] order the next all-to-all sends — no transposition pass;
    // tw2    : the last pass also multiplies element idx of member b by w^((tw2_row0 + b) * idx), w the 2^tw2_L-th root whose
    //          power table is tw2 — the twiddle step of the four-step decomposition, fused into the store.
    int in_map, out_map;
    unsigned in_s, out_s;
    unsigned long long in_A, in_B, out_A, out_B;
    const uint4* tw2;
    unsigned tw2_L;
    unsigned long long tw2_row0;
    // out_map == 2: the transposed store goes STRAIGHT INTO THE PEERS' receive buffers over NVLink (CUDA IPC pointers): result
    // idx of member b belongs to rank h = idx >> peer_rows_log and lands at peer[h][peer_self_off + (idx & mask) * batch + b]
    // — the exchange of the distributed NTT is the transform's own last store, tile by tile, with no separate collective.
    uint4* peer[8];
    unsigned peer_rows_log;
    unsigned long long peer_self_off;
};

__device__ __forceinline__ Fr ld_fr(const uint4* p, size_t i) {
    uint4 a = p[2 * i], b = p[2 * i + 1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ Fr ldg_fr(const uint4* p, size_t i) {
    uint4 a = __ldg(p + 2 * i), b = __ldg(p + 2 * i + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(uint4* p, size_t i, const Fr& v) {
    p[2 * i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[2 * i + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

__device__ __forceinline__ Fr sm_ld(const uint4* slo, const uint4* shi, uint32_t e) {
    uint4 l = slo[e], h = shi[e];
    Fr v;
    v.l[0] = l.x; v.l[1] = l.y; v.l[2] = l.z; v.l[3] = l.w; v.l[4] = h.x; v.l[5] = h.y; v.l[6] = h.z; v.l[7] = h.w;
    return v;
}
__device__ __forceinline__ void sm_st(uint4* slo, uint4* shi, uint32_t e, const Fr& v) {
    slo[e] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    shi[e] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// (x, y) <- (x + w y, x - w y)
__device__ __forceinline__ void butterfly_w(Fr& x, Fr& y, const Fr& w) {
    y = fp_mul<FrP>(y, w);
    Fr u = fp_add<FrP>(x, y);
    y = fp_sub<FrP>(x, y);
    x = u;
}
// w = 1 (the reference skips that multiply too, arithmetic.rs:213-219)
__device__ __forceinline__ void butterfly_1(Fr& x, Fr& y) {
    Fr u = fp_add<FrP>(x, y);
    y = fp_sub<FrP>(x, y);
    x = u;
}

// One pass = up to 8 butterfly stages on a tile of T = 2^(r+q) elements in shared memory. T/4 threads; each thread
// carries FOUR elements through TWO stages in registers (a radix-4 step = 2 + 2 butterflies), so a pass needs r/2
// shared-memory round trips and barriers instead of r, and every thread has two independent multiplications in flight.
__global__ void __launch_bounds__(256, NTT_MIN_CTAS) ntt_pass_kernel(const __grid_constant__ NttPassArgs a) {
    extern __shared__ uint4 sm[];
    const uint32_t T = 1u << (a.r + a.q);
    uint4* slo = sm;
    uint4* shi = sm + T;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const uint32_t blk = blockIdx.x;
    const size_t boff = (size_t)blockIdx.y << a.L;  // batched transforms: member blockIdx.y of gridDim.y independent vectors
    const uint32_t qmask = (1u << a.q) - 1u;
    // bits of the global index owned by this CTA
    uint32_t lo = 0, hi = 0;
    if (!a.first) {
        lo = blk & ((1u << (a.s0 - a.q)) - 1u);
        hi = blk >> (a.s0 - a.q);
    }
    // ---- load ----------------------------------------------------------------------------------------------------
    for (uint32_t e = tid; e < T; e += nthr) {
        uint32_t row = e >> a.q, col = e & qmask;
        Fr v;
        if (a.first) {
            // destination (bit-reversed-order) index i; natural source index j = bitrev_L(i)
            uint32_t i = row | ((blk | (col << (a.L - a.r - a.q))) << a.r);
            uint32_t j = __brev(i) >> (32 - a.L);
            if ((unsigned long long)j < a.n_in) {
                v = ld_fr(a.src, a.in_map ? (size_t)((j >> a.in_s) * a.in_A + blockIdx.y * a.in_B + (j & ((1u << a.in_s) - 1u))) : boff + j);
                if (a.pre_mode == 1) {
                    uint32_t m = j % 3u;
                    if (m) v = fp_mul<FrP>(v, a.pre[m]);
                } else if (a.pre_mode == 2) {
                    v = fp_mul<FrP>(v, a.pre[j & (uint32_t)(a.pre_len - 1)]);
                }
            } else {
                v = Fr::zero();
            }
        } else {
            size_t i = (size_t)col | ((size_t)lo << a.q) | ((size_t)row << a.s0) | ((size_t)hi << (a.s0 + a.r));
            v = ld_fr(a.src, boff + i);
        }
        sm_st(slo, shi, e, v);
    }
    __syncthreads();
    // ---- butterflies: stage s = s0 + t pairs rows that differ in bit t ----------------------------------------------
    // twiddle index of a butterfly whose upper row is `row` at stage t: (i mod 2^s) << (L-1-s), s = s0 + t
    //   (reference arithmetic.rs:263-272: twiddles[(i + 1) * twiddle_chunk] over the chunk-local index)
    // The twiddles of a radix-4 step are fetched (L2, ~700 cycles) BEFORE the shared-memory loads and the first multiply, all
    // three at once: behind a per-butterfly `if (twi != 0)` the loads stayed where they were and every butterfly pair waited
    // for its own round trip. W[0] = 1 in Montgomery form, so an index of 0 needs no special case (same limbs); only the one
    // step whose twiddles are 1 for EVERY lane — stages 1-2 of the first pass: one multiplication instead of four — is
    // specialised, on a warp-uniform condition.
    int t = 0;
    for (; t + 1 < a.r; t += 2) {
        const bool first_step = a.first && t == 0;
        for (uint32_t qd = tid; qd < (T >> 2); qd += nthr) {
            const uint32_t c = qd & qmask, rq = qd >> a.q;
            const uint32_t jlow = a.first ? 0u : (c | (lo << a.q));
            const uint32_t r00 = ((rq >> t) << (t + 2)) | (rq & ((1u << t) - 1u));
            const uint32_t r01 = r00 | (1u << t), r10 = r00 | (2u << t), r11 = r00 | (3u << t);
            const uint32_t e00 = (r00 << a.q) | c, e01 = (r01 << a.q) | c, e10 = (r10 << a.q) | c, e11 = (r11 << a.q) | c;
            const uint32_t low_t = r00 & ((1u << t) - 1u);
            const int s = a.s0 + t;
            const uint32_t tw0 = (jlow | (low_t << a.s0)) << (a.L - 1 - s);              // stage t (same for both pairs)
            const uint32_t tw1a = (jlow | (low_t << a.s0)) << (a.L - 2 - s);              // stage t+1, rows r00 / r10
            const uint32_t tw1b = (jlow | ((low_t | (1u << t)) << a.s0)) << (a.L - 2 - s);  // stage t+1, rows r01 / r11
            const Fr w1b = ldg_fr(a.tw, tw1b);
            if (first_step) {  // tw0 == tw1a == 0 on every lane
                Fr x00 = sm_ld(slo, shi, e00), x01 = sm_ld(slo, shi, e01), x10 = sm_ld(slo, shi, e10), x11 = sm_ld(slo, shi, e11);
                butterfly_1(x00, x01);
                butterfly_1(x10, x11);
                butterfly_1(x00, x10);
                butterfly_w(x01, x11, w1b);
                sm_st(slo, shi, e00, x00); sm_st(slo, shi, e01, x01); sm_st(slo, shi, e10, x10); sm_st(slo, shi, e11, x11);
            } else {
                const Fr w0 = ldg_fr(a.tw, tw0), w1a = ldg_fr(a.tw, tw1a);
                Fr x00 = sm_ld(slo, shi, e00), x01 = sm_ld(slo, shi, e01), x10 = sm_ld(slo, shi, e10), x11 = sm_ld(slo, shi, e11);
                butterfly_w(x00, x01, w0);
                butterfly_w(x10, x11, w0);
                butterfly_w(x00, x10, w1a);
                butterfly_w(x01, x11, w1b);
                sm_st(slo, shi, e00, x00); sm_st(slo, shi, e01, x01); sm_st(slo, shi, e10, x10); sm_st(slo, shi, e11, x11);
            }
        }
        __syncthreads();
    }
    if (t < a.r) {  // odd number of stages: one plain radix-2 stage
        const bool first_stage = a.first && t == 0;  // a one-stage first pass (tiny transforms): every twiddle is 1
        for (uint32_t b = tid; b < (T >> 1); b += nthr) {
            const uint32_t c = b & qmask, rb = b >> a.q;
            const uint32_t jlow = a.first ? 0u : (c | (lo << a.q));
            const uint32_t r0 = ((rb >> t) << (t + 1)) | (rb & ((1u << t) - 1u));
            const uint32_t r1 = r0 | (1u << t);
            const uint32_t e0 = (r0 << a.q) | c, e1 = (r1 << a.q) | c;
            const int s = a.s0 + t;
            const uint32_t twi = (jlow | ((r0 & ((1u << t) - 1u)) << a.s0)) << (a.L - 1 - s);
            Fr x = sm_ld(slo, shi, e0), y = sm_ld(slo, shi, e1);
            if (first_stage) butterfly_1(x, y);
            else butterfly_w(x, y, ldg_fr(a.tw, twi));
            sm_st(slo, shi, e0, x); sm_st(slo, shi, e1, y);
        }
        __syncthreads();
    }
    // ---- store ---------------------------------------------------------------------------------------------------
    for (uint32_t e = tid; e < T; e += nthr) {
        uint32_t row = e >> a.q, col = e & qmask;
        size_t i;
        if (a.first) i = (size_t)row | ((size_t)(blk | (col << (a.L - a.r - a.q))) << a.r);
        else i = (size_t)col | ((size_t)lo << a.q) | ((size_t)row << a.s0) | ((size_t)hi << (a.s0 + a.r));
        Fr v = sm_ld(slo, shi, e);
        if (a.post_mode) v = fp_mul<FrP>(v, a.post[a.post_mode == 1 ? 0 : (int)(i % 3)]);
        if (a.tw2) {
            const size_t ex = ((a.tw2_row0 + blockIdx.y) * i) & (((size_t)1 << a.tw2_L) - 1);
            const size_t half = (size_t)1 << (a.tw2_L - 1);
            if (ex) v = fp_mul<FrP>(v, ex < half ? ldg_fr(a.tw2, ex) : fp_neg<FrP>(ldg_fr(a.tw2, ex - half)));
        }
        if (a.out_map == 2) {
            const size_t h = i >> a.peer_rows_log;
            st_fr(a.peer[h], (size_t)(a.peer_self_off + (i & (((size_t)1 << a.peer_rows_log) - 1)) * gridDim.y + blockIdx.y), v);
        } else {
            st_fr(a.dst, a.out_map ? (size_t)((i >> a.out_s) * a.out_A + blockIdx.y * a.out_B + (i & (((size_t)1 << a.out_s) - 1))) : boff + i, v);
        }
    }
}

// W[i] = omega^i for i < count. pw[t] = omega^(2^t). Each thread owns TW_RUN consecutive entries.
constexpr int TW_RUN = 32;
struct TwArgs {
    uint4* out;
    unsigned long long count;
    Fr pw[28];
};
__global__ void __launch_bounds__(256) ntt_twiddle_kernel(const __grid_constant__ TwArgs a) {
    unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long i0 = t * TW_RUN;
    if (i0 >= a.count) return;
    Fr x = Fr::one();
    for (int b = 0; b < 28; b++)
        if ((i0 >> b) & 1ull) x = fp_mul<FrP>(x, a.pw[b]);
    for (int k = 0; k < TW_RUN && i0 + k < a.count; k++) {
        st_fr(a.out, i0 + k, x);
        x = fp_mul<FrP>(x, a.pw[0]);
    }
}

// element-wise x[i] *= tab[i & (len-1)] (only used when a vanishing-division table is too long for the fused path)
__global__ void fr_scale_table_kernel(uint4* a, size_t n, const uint4* tab, uint32_t len) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = fp_mul<FrP>(ld_fr(a, i), ldg_fr(tab, i & (len - 1)));
    st_fr(a, i, v);
}
// n == 1 corner: a[0] = a[0] * f
__global__ void fr_scale_one_kernel(const uint4* src, uint4* dst, Fr f) {
    Fr v = fp_mul<FrP>(ld_fr(src, 0), f);
    st_fr(dst, 0, v);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
struct TwKey {
    uint64_t w[4];
    uint32_t L;
    bool operator<(const TwKey& o) const {
        if (L != o.L) return L < o.L;
        for (int i = 0; i < 4; i++)
            if (w[i] != o.w[i]) return w[i] < o.w[i];
        return false;
    }
};
struct TwEntry { void* d; size_t bytes; unsigned long long last_use; };
static std::map<TwKey, TwEntry> g_tw;
static unsigned long long g_tw_clock = 0;
static size_t g_tw_bytes = 0;
static const size_t TW_CACHE_LIMIT = (size_t)24 << 30;  // 24 GiB of the 180 GB HBM for twiddle tables
static Scratch g_ntt_scratch;

Fr fr_from_u64x4(const uint64_t* p) {
    Fr r;
    for (int i = 0; i < 4; i++) { r.l[2 * i] = (uint32_t)p[i]; r.l[2 * i + 1] = (uint32_t)(p[i] >> 32); }
    return r;
}

void ntt_release_all() {
    for (auto& kv : g_tw) cudaFree(kv.second.d);
    g_tw.clear();
    g_tw_bytes = 0;
    g_ntt_scratch.release();
}

static int get_twiddles(const uint64_t omega[4], uint32_t L, const uint4** out) {
    TwKey key;
    memcpy(key.w, omega, 32);
    key.L = L;
    auto it = g_tw.find(key);
    if (it != g_tw.end()) {
        it->second.last_use = ++g_tw_clock;
        *out = (const uint4*)it->second.d;
        return 0;
    }
    size_t count = L >= 1 ? ((size_t)1 << (L - 1)) : 1;
    size_t bytes = count * 32;
    while (g_tw_bytes + bytes > TW_CACHE_LIMIT && !g_tw.empty()) {  // evict least recently used
        auto victim = g_tw.begin();
        for (auto i2 = g_tw.begin(); i2 != g_tw.end(); ++i2)
            if (i2->second.last_use < victim->second.last_use) victim = i2;
        CQB_CUDA(cudaStreamSynchronize(ctx().stream));
        cudaFree(victim->second.d);
        g_tw_bytes -= victim->second.bytes;
        g_tw.erase(victim);
    }
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return fail(CQB_E_OOM, "twiddle table: cudaMalloc(%zu) failed", bytes);
    TwArgs ta;
    ta.out = (uint4*)d;
    ta.count = count;
    Fr w = fr_from_u64x4(omega);
    for (int b = 0; b < 28; b++) { ta.pw[b] = w; w = fp_sqr<FrP>(w); }  // host path of fp.cuh
    size_t threads = (count + TW_RUN - 1) / TW_RUN;
    unsigned grid = (unsigned)((threads + 255) / 256);
    ntt_twiddle_kernel<<<grid, 256, 0, ctx().stream>>>(ta);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    g_tw[key] = TwEntry{d, bytes, ++g_tw_clock};
    g_tw_bytes += bytes;
    *out = (const uint4*)d;
    return 0;
}

int ntt_get_twiddles(const uint64_t omega[4], uint32_t L, const void** out) {
    const uint4* p = nullptr;
    CQB_TRY(get_twiddles(omega, L, &p));
    *out = p;
    return 0;
}

// out[i] = base^i for i < count (the twiddle kernel on an arbitrary base); used by the SRS generator
int fr_powers_run(const uint64_t base[4], size_t count, void* d_out) {
    if (count == 0) return 0;
    TwArgs ta;
    ta.out = (uint4*)d_out;
    ta.count = count;
    Fr w = fr_from_u64x4(base);
    for (int b = 0; b < 28; b++) { ta.pw[b] = w; w = fp_sqr<FrP>(w); }
    size_t threads = (count + TW_RUN - 1) / TW_RUN;
    ntt_twiddle_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx().stream>>>(ta);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

// d_src may alias d_dst (in place). All pointers are device pointers to 32-byte Fr elements.
int ntt_run(const void* d_src, void* d_dst, uint32_t L, const uint64_t omega[4], const NttFused& f, uint32_t batch) {
    if (batch == 0) return 0;
    if (batch > 65535u) return fail(CQB_E_BAD_SIZE, "NTT batch of %u exceeds 65535", batch);
    if (L > 28) return fail(CQB_E_BAD_SIZE, "log_n = %u exceeds Fr::S = 28 (bn256/fr.rs:72)", L);
    cudaStream_t st = ctx().stream;
    size_t n = (size_t)1 << L;
    static bool attr_set = false;
    if (!attr_set) {
        CQB_CUDA(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set = true;
    }
    if (L == 0) {
        Fr m = Fr::one();
        if (f.pre_mode == 2) m = fp_mul<FrP>(m, f.pre[0]);
        if (f.post_mode) m = fp_mul<FrP>(m, f.post[0]);
        if (batch != 1) return fail(CQB_E_BAD_SIZE, "batched NTT of length 1 is not supported");
        fr_scale_one_kernel<<<1, 1, 0, st>>>((const uint4*)d_src, (uint4*)d_dst, m);
        CQB_LAUNCHED();
        CQB_CUDA(cudaGetLastError());
        return 0;
    }
    const uint4* tw = nullptr;
    CQB_TRY(get_twiddles(omega, L, &tw));
    int P = (int)((L + NTT_MAX_R - 1) / NTT_MAX_R);
    int rs[8];
    for (int p = 0; p < P; p++) rs[p] = (int)L / P + (p < (int)L % P ? 1 : 0);
    bool inplace = (d_src == d_dst);
    if (inplace && (f.in_map || f.out_map)) return fail(CQB_E_BAD_ARG, "mapped NTT must be out of place");
    const bool use_scratch = inplace || f.out_map;  // a mapped (scattered) last store must not land on data still to be read
    uint4* work = (uint4*)d_dst;
    if (use_scratch) {
        CQB_TRY(g_ntt_scratch.ensure((size_t)batch * n * 32));
        work = g_ntt_scratch.as<uint4>();
    }
    int s0 = 0;
    for (int p = 0; p < P; p++) {
        NttPassArgs a;
        a.L = (int)L;
        a.s0 = s0;
        a.r = rs[p];
        a.first = (p == 0);
        int qmax = (p == 0) ? ((int)L - a.r) : s0;
        a.q = NTT_TILE_LOG - a.r;
        if (a.q > qmax) a.q = qmax;
        a.src = (p == 0) ? (const uint4*)d_src : work;
        a.dst = (p == P - 1 && use_scratch && (P > 1 || !inplace)) ? (uint4*)d_dst : work;
        a.tw = tw;
        a.n_in = f.n_in ? f.n_in : n;
        a.pre_mode = (p == 0) ? f.pre_mode : 0;
        a.pre_len = f.pre_len;
        for (int i = 0; i < NTT_PRE_MAX; i++) a.pre[i] = f.pre[i];
        a.post_mode = (p == P - 1) ? f.post_mode : 0;
        a.in_map = (p == 0) ? f.in_map : 0;
        a.in_s = f.in_s; a.in_A = f.in_A; a.in_B = f.in_B;
        a.out_map = (p == P - 1) ? f.out_map : 0;
        a.out_s = f.out_s; a.out_A = f.out_A; a.out_B = f.out_B;
        for (int h = 0; h < 8; h++) a.peer[h] = (uint4*)f.peer[h];
        a.peer_rows_log = f.peer_rows_log;
        a.peer_self_off = f.peer_self_off;
        a.tw2 = (p == P - 1) ? (const uint4*)f.tw2 : nullptr;
        a.tw2_L = f.tw2_L; a.tw2_row0 = f.tw2_row0;
        for (int i = 0; i < 3; i++) a.post[i] = f.post[i];
        int T = 1 << (a.r + a.q);
        unsigned grid = (unsigned)(n >> (a.r + a.q));
        int threads = T >> 2;
        if (threads < 1) threads = 1;
        ntt_pass_kernel<<<dim3(grid, batch), threads, (size_t)T * 32, st>>>(a);
        CQB_LAUNCHED();
        CQB_CUDA(cudaGetLastError());
        s0 += a.r;
    }
    if (inplace && P == 1) CQB_CUDA(cudaMemcpyAsync(d_dst, work, (size_t)batch * n * 32, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ---- pieces of the distributed four-step NTT (sharded.py ShardedNTT) ------------------------------------------------------
// a[r][c] *= omega^((row0 + r) * c), omega = the 2^L-th root whose table tw[i] = omega^i (i < 2^(L-1)) is resident:
// the twiddle step between the column and the row transforms of the four-step decomposition
__global__ void __launch_bounds__(256) fr_mul_omega_powers_kernel(uint4* __restrict__ a, size_t rows, size_t cols, size_t row0,
                                                                  const uint4* __restrict__ tw, uint32_t L) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const size_t r = i / cols, c = i % cols;
    const size_t e = ((row0 + r) * c) & (((size_t)1 << L) - 1);
    const size_t half = (size_t)1 << (L - 1);
    if (e == 0) return;
    Fr w = e < half ? ldg_fr(tw, e) : fp_neg<FrP>(ldg_fr(tw, e - half));  // omega^(n/2) = -1
    st_fr(a, i, fp_mul<FrP>(ld_fr(a, i), w));
}
int fr_mul_omega_powers_run(void* d_a, size_t rows, size_t cols, size_t row0, const uint64_t omega[4], uint32_t L) {
    if (rows == 0 || cols == 0) return 0;
    if (L == 0 || L > 28) return fail(CQB_E_BAD_SIZE, "omega powers: log_n = %u out of range", L);
    const uint4* tw = nullptr;
    CQB_TRY(get_twiddles(omega, L, &tw));
    size_t total = rows * cols;
    fr_mul_omega_powers_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx().stream>>>((uint4*)d_a, rows, cols, row0, tw, L);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}
// out[c][r] = in[r][c] for 32-byte elements: 32 x 32 tiles staged in shared memory as two uint4 planes (+1 padding), so that
// both the global reads and the global writes are 128-bit and row-contiguous
__global__ void __launch_bounds__(256) fr_transpose_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint32_t rows, uint32_t cols) {
    __shared__ uint4 tlo[32][33], thi[32][33];
    const uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (uint32_t j = ty; j < 32; j += 8) {
        const uint32_t r = r0 + j, c = c0 + tx;
        if (r < rows && c < cols) {
            const size_t i = (size_t)r * cols + c;
            tlo[j][tx] = in[2 * i];
            thi[j][tx] = in[2 * i + 1];
        }
    }
    __syncthreads();
    for (uint32_t j = ty; j < 32; j += 8) {
        const uint32_t c = c0 + j, r = r0 + tx;
        if (r < rows && c < cols) {
            const size_t o = (size_t)c * rows + r;
            out[2 * o] = tlo[tx][j];
            out[2 * o + 1] = thi[tx][j];
        }
    }
}
int fr_transpose_run(const void* d_in, void* d_out, size_t rows, size_t cols) {
    if (rows == 0 || cols == 0) return 0;
    if (rows > 0xffffffffull || cols > 0xffffffffull) return fail(CQB_E_BAD_SIZE, "transpose dimensions exceed 32 bits");
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    if (grid.y > 65535u) return fail(CQB_E_BAD_SIZE, "transpose: more than 2^21 rows");
    fr_transpose_kernel<<<grid, 256, 0, ctx().stream>>>((const uint4*)d_in, (uint4*)d_out, (uint32_t)rows, (uint32_t)cols);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

int fr_scale_table(void* d_a, size_t n, const void* d_tab, uint32_t len) {
    fr_scale_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>((uint4*)d_a, n, (const uint4*)d_tab, len);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cqb

// srs.cu — SURVEY.md §8(f) row 3: KZG SRS generation on the device, replacing the parallelize()d scalar-multiplication
// loops of ParamsKZG::setup_from_toxic_waste / setup (reference halo2_proofs/src/poly/kzg/commitment.rs:209-276, 280-348)
// and the G1 part of TableSRS::setup_from_toxic_waste (:73-141):
//     g[i]          = [s^i] G
//     g_lagrange[i] = [ (s^n - 1)/n * w^i / (s - w^i) ] G        w = the n-th root of unity (commitment.rs:234-250)
// The reference walks `current_g *= s` (a 256-step double-and-add per point) inside rayon chunks and normalises with
// batch_normalize; only the affine normal forms are observable, so the device computes the same points as
//   1. scalar vectors s^i, w^i (parallel powers), d_i = s - w^i, batch inversion (Montgomery's trick per thread run,
//      zeros skipped like ff::BatchInvert), l_i = mult * w^i * d_i^-1;
//   2. fixed-base multiplication by the generator: a resident table T[w][d] = d * 2^(16 w) G (16 windows x 2^15 signed
//      digits x 64 B = 32 MiB, L2-resident), 16 table lookups + XYZZ mixed adds per point, one shared inversion per run
//      of 32 points (batch_normalize, derive/curve.rs:362-397).
// Also exposes the element-wise helpers this needs, which are "next" rows of their own (§8f row 4): batch inversion and
// power vectors.
#include "internal.h"

namespace cqb {

__device__ __forceinline__ Fq s_ld_fq(const uint4* p) {
    uint4 a = p[0], b = p[1];
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void s_st_fq(uint4* p, const Fq& v) {
    p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fr s_ld_fr(const uint4* p) {
    uint4 a = p[0], b = p[1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void s_st_fr(uint4* p, const Fr& v) {
    p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// ---- Fr batch inversion in place: a[i] <- a[i]^-1, zeros stay zero (ff::BatchInvert semantics) ------------------------
// One thread inverts a run of `run` consecutive elements with Montgomery's trick: 3 multiplications per element plus one
// field inversion per run. The run length grows with n (32 ... 256): long runs amortise the inversion, short ones keep
// enough threads in flight for small vectors. (The binary-Euclid inversion was tried here too: never faster — the lanes of
// a warp run different trip counts.)
__global__ void __launch_bounds__(128) fr_batch_invert_kernel(uint4* __restrict__ a, size_t n, uint4* __restrict__ tmp, uint32_t run) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t p0 = t * run;
    if (p0 >= n) return;
    size_t cnt = (n - p0 < (size_t)run) ? (n - p0) : (size_t)run;
    Fr prod = Fr::one();
    for (size_t j = 0; j < cnt; j++) {
        Fr v = s_ld_fr(a + (p0 + j) * 2);
        s_st_fr(tmp + (p0 + j) * 2, prod);
        if (!v.is_zero()) prod = fp_mul<FrP>(prod, v);
    }
    Fr inv = fp_inv<FrP>(prod);
    for (size_t j = cnt; j-- > 0;) {
        Fr v = s_ld_fr(a + (p0 + j) * 2);
        if (v.is_zero()) continue;
        Fr pre = s_ld_fr(tmp + (p0 + j) * 2);
        Fr vi = fp_mul<FrP>(inv, pre);
        inv = fp_mul<FrP>(inv, v);
        s_st_fr(a + (p0 + j) * 2, vi);
    }
}

// d[i] = s - w[i]
__global__ void srs_diff_kernel(const uint4* __restrict__ w, size_t n, Fr s, uint4* __restrict__ d) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = fp_sub<FrP>(s, s_ld_fr(w + 2 * i));
    s_st_fr(d + 2 * i, v);
}
// l[i] = mult * w[i] * dinv[i]   (reference commitment.rs:247: multiplier * root_pow * (s - root_pow).invert())
__global__ void srs_lagrange_scalar_kernel(const uint4* __restrict__ w, uint4* __restrict__ dinv_inout, size_t n, Fr mult) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = fp_mul<FrP>(fp_mul<FrP>(mult, s_ld_fr(w + 2 * i)), s_ld_fr(dinv_inout + 2 * i));
    s_st_fr(dinv_inout + 2 * i, v);
}

// o[i] = l[i] * winv[i] - c   (reference commitment.rs:146-168: l_i * w^-i - [x^(N-1)]/N, as a scalar of G)
__global__ void srs_opening_scalar_kernel(const uint4* __restrict__ l, uint4* __restrict__ winv_inout, size_t n, Fr c) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = fp_sub<FrP>(fp_mul<FrP>(s_ld_fr(l + 2 * i), s_ld_fr(winv_inout + 2 * i)), c);
    s_st_fr(winv_inout + 2 * i, v);
}

// ---- fixed-base table: T[w][d-1] = d * B_w, B_w = 2^(16 w) G, d = 1..32768 --------------------------------------------
constexpr int FB_C = 16, FB_WIN = 16, FB_NB = 1 << (FB_C - 1), FB_RUN = 64;
__global__ void __launch_bounds__(128) fb_table_kernel(const uint4* __restrict__ rows, uint4* __restrict__ table, uint4* __restrict__ tmp) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_win = FB_NB / FB_RUN;
    if (t >= (size_t)FB_WIN * per_win) return;
    uint32_t w = (uint32_t)(t / per_win), chunk = (uint32_t)(t % per_win);
    Fq bx = s_ld_fq(rows + (size_t)w * 4), by = s_ld_fq(rows + (size_t)w * 4 + 2);
    uint32_t m0 = chunk * FB_RUN + 1;  // first multiple of this run
    G1Xyzz P = G1Xyzz::identity();
    for (int b = 31 - __clz(m0); b >= 0; b--) {
        P = g1_double(P);
        if ((m0 >> b) & 1u) g1_madd(P, bx, by);
    }
    uint4* tslot = tmp + t * FB_RUN * 10;
    Fq prod = Fq::one();
    for (int j = 0; j < FB_RUN; j++) {
        uint4* slot = tslot + j * 10;
        s_st_fq(slot, P.x); s_st_fq(slot + 2, P.y); s_st_fq(slot + 4, P.zz); s_st_fq(slot + 6, P.zzz);
        s_st_fq(slot + 8, prod);
        prod = fp_mul<FqP>(prod, fp_mul<FqP>(P.zz, P.zzz));
        g1_madd(P, bx, by);
    }
    Fq inv = fp_inv<FqP>(prod);
    for (int j = FB_RUN - 1; j >= 0; j--) {
        const uint4* slot = tslot + j * 10;
        Fq x = s_ld_fq(slot), y = s_ld_fq(slot + 2), zz = s_ld_fq(slot + 4), zzz = s_ld_fq(slot + 6), pre = s_ld_fq(slot + 8);
        Fq zi = fp_mul<FqP>(inv, pre);
        inv = fp_mul<FqP>(inv, fp_mul<FqP>(zz, zzz));
        uint4* dst = table + ((size_t)w * FB_NB + (m0 - 1) + j) * 4;
        s_st_fq(dst, fp_mul<FqP>(x, fp_mul<FqP>(zi, zzz)));
        s_st_fq(dst + 2, fp_mul<FqP>(y, fp_mul<FqP>(zi, zz)));
    }
}

// out[i] = [k_i] G : 16 signed 16-bit digits -> 16 table lookups + mixed adds; one shared inversion per run of points
constexpr int FBM_RUN = 32;
__global__ void __launch_bounds__(128) fb_mul_kernel(const uint4* __restrict__ scalars, size_t n, const uint4* __restrict__ table,
                                                     uint4* __restrict__ tmp, uint4* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t p0 = t * FBM_RUN;
    if (p0 >= n) return;
    size_t cnt = (n - p0 < (size_t)FBM_RUN) ? (n - p0) : (size_t)FBM_RUN;
    Fq prod = Fq::one();
    for (size_t j = 0; j < cnt; j++) {
        Fr k = fp_from_mont<FrP>(s_ld_fr(scalars + (p0 + j) * 2));
        G1Xyzz acc = G1Xyzz::identity();
        uint32_t carry = 0;
#pragma unroll 1
        for (int w = 0; w < FB_WIN; w++) {
            uint32_t d = ((k.l[w >> 1] >> ((w & 1) * 16)) & 0xffffu) + carry;
            carry = 0;
            uint32_t neg = 0;
            if (d > (uint32_t)FB_NB) { d = (1u << FB_C) - d; carry = 1; neg = 1; }
            if (d) {
                const uint4* e = table + ((size_t)w * FB_NB + (d - 1)) * 4;
                uint4 a0 = __ldg(e), a1 = __ldg(e + 1), a2 = __ldg(e + 2), a3 = __ldg(e + 3);
                Fq x, y;
                x.l[0] = a0.x; x.l[1] = a0.y; x.l[2] = a0.z; x.l[3] = a0.w; x.l[4] = a1.x; x.l[5] = a1.y; x.l[6] = a1.z; x.l[7] = a1.w;
                y.l[0] = a2.x; y.l[1] = a2.y; y.l[2] = a2.z; y.l[3] = a2.w; y.l[4] = a3.x; y.l[5] = a3.y; y.l[6] = a3.z; y.l[7] = a3.w;
                if (neg) y = fp_neg<FqP>(y);
                g1_madd(acc, x, y);
            }
        }
        // scalars are < r < 2^254, so the top window (14 bits + carry) never overflows: no carry is left here
        uint4* slot = tmp + (p0 + j) * 10;
        s_st_fq(slot, acc.x); s_st_fq(slot + 2, acc.y); s_st_fq(slot + 4, acc.zz); s_st_fq(slot + 6, acc.zzz);
        s_st_fq(slot + 8, prod);
        if (!acc.is_identity()) prod = fp_mul<FqP>(prod, fp_mul<FqP>(acc.zz, acc.zzz));
    }
    Fq inv = fp_inv<FqP>(prod);
    for (size_t j = cnt; j-- > 0;) {
        const uint4* slot = tmp + (p0 + j) * 10;
        Fq x = s_ld_fq(slot), y = s_ld_fq(slot + 2), zz = s_ld_fq(slot + 4), zzz = s_ld_fq(slot + 6), pre = s_ld_fq(slot + 8);
        Fq ax = Fq::zero(), ay = Fq::zero();
        if (!zz.is_zero()) {
            Fq zi = fp_mul<FqP>(inv, pre);
            inv = fp_mul<FqP>(inv, fp_mul<FqP>(zz, zzz));
            ax = fp_mul<FqP>(x, fp_mul<FqP>(zi, zzz));
            ay = fp_mul<FqP>(y, fp_mul<FqP>(zi, zz));
        }
        uint4* dst = out + (p0 + j) * 4;
        s_st_fq(dst, ax);
        s_st_fq(dst + 2, ay);
    }
}

// -------------------------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------------------------
static void* g_fb_table = nullptr;  // 16 x 32768 x 64 B, built on first use
static Scratch g_srs_tmp, g_srs_vec;

void srs_release_all() {
    if (g_fb_table) cudaFree(g_fb_table);
    g_fb_table = nullptr;
    g_srs_tmp.release();
    g_srs_vec.release();
}

int fr_batch_invert_run(void* d_a, size_t n) {
    if (n == 0) return 0;
    CQB_TRY(g_srs_tmp.ensure(n * 32));
    uint32_t run = 32;
    while (run < 256 && n / (run * 2) >= 32768) run *= 2;  // keep >= 32k threads in flight, then lengthen the runs
    // measured (B200, ms at 2^20 / 2^22 / 2^24): run 32: 0.31 / 0.96 / 3.62, 64: 0.30 / 0.73 / 2.27, 128: 0.44 / 0.58 / 1.65,
    // 256: 0.73 / 0.77 / 1.59; the binary inversion was not faster at any size (divergent trip counts)
    size_t threads = (n + run - 1) / run;
    fr_batch_invert_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx().stream>>>((uint4*)d_a, n, g_srs_tmp.as<uint4>(), run);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

static int ensure_fb_table() {
    if (g_fb_table) return 0;
    cudaStream_t st = ctx().stream;
    // rows B_w = 2^(16 w) G through the MSM precompute kernel on the single point G
    Fq gx = Fq::one(), gy = fp_dbl<FqP>(Fq::one());  // G = (1, 2), bn256/curve.rs:66-67
    uint32_t gpt[16];
    for (int i = 0; i < 8; i++) { gpt[i] = gx.l[i]; gpt[8 + i] = gy.l[i]; }
    void* d_rows = nullptr;
    if (cudaMalloc(&d_rows, (size_t)FB_WIN * 64 + 64) != cudaSuccess) return fail(CQB_E_OOM, "fixed-base rows: cudaMalloc failed");
    void* d_g = (char*)d_rows + (size_t)FB_WIN * 64;
    CQB_CUDA(cudaMemcpyAsync(d_g, gpt, 64, cudaMemcpyHostToDevice, st));
    CQB_CUDA(cudaStreamSynchronize(st));  // gpt is a stack buffer
    CQB_TRY(msm_precompute_table(d_g, 1, FB_C, d_rows));  // nwin(16) = 16 rows
    size_t table_bytes = (size_t)FB_WIN * FB_NB * 64;
    if (cudaMalloc(&g_fb_table, table_bytes) != cudaSuccess) { cudaFree(d_rows); g_fb_table = nullptr; return fail(CQB_E_OOM, "fixed-base table: cudaMalloc failed"); }
    size_t threads = (size_t)FB_WIN * (FB_NB / FB_RUN);
    CQB_TRY(g_srs_tmp.ensure(threads * FB_RUN * 160));
    fb_table_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>((const uint4*)d_rows, (uint4*)g_fb_table, g_srs_tmp.as<uint4>());
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    CQB_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_rows);
    return 0;
}

int g1_generator_mul_run(const void* d_scalars, size_t n, void* d_out) {
    if (n == 0) return 0;
    CQB_TRY(ensure_fb_table());
    cudaStream_t st = ctx().stream;
    const size_t CHUNK = (size_t)1 << 22;
    CQB_TRY(g_srs_tmp.ensure(std::min(n, CHUNK) * 160));
    for (size_t off = 0; off < n; off += CHUNK) {
        size_t m = std::min(CHUNK, n - off);
        size_t threads = (m + FBM_RUN - 1) / FBM_RUN;
        fb_mul_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>((const uint4*)d_scalars + off * 2, m, (const uint4*)g_fb_table,
                                                                          g_srs_tmp.as<uint4>(), (uint4*)d_out + off * 4);
        CQB_LAUNCHED();
    }
    CQB_CUDA(cudaGetLastError());
    return 0;
}

// commitment.rs:209-276: g and g_lagrange for n = 2^k from the toxic waste s (Montgomery limbs); with d_opening_at_0 also
// TableSRS's g_lagrange_opening_at_0 (commitment.rs:143-170): [(L_i(x) - L_i(0))/x]_1 = w^-i [L_i(x)]_1 - (1/N) [x^(N-1)]_1
int srs_setup_run(uint32_t k, const uint64_t s_limbs[4], void* d_g, void* d_g_lagrange, void* d_opening_at_0) {
    if (k > 28) return fail(CQB_E_BAD_SIZE, "k = %u exceeds Fr::S = 28 (commitment.rs:212)", k);
    cudaStream_t st = ctx().stream;
    size_t n = (size_t)1 << k;
    Fr s = fr_from_u64x4(s_limbs);
    CQB_TRY(g_srs_vec.ensure(n * 32 * 3));
    void* d_pow = g_srs_vec.p;
    void* d_aux = (char*)g_srs_vec.p + n * 32;
    void* d_aux2 = (char*)g_srs_vec.p + 2 * n * 32;
    // g[i] = [s^i] G
    CQB_TRY(fr_powers_run(s_limbs, n, d_pow));
    CQB_TRY(g1_generator_mul_run(d_pow, n, d_g));
    // root = ROOT_OF_UNITY_INV.invert() squared (S - k) times (commitment.rs:235-238); n_inv; multiplier (:239-241)
    Fr root;
    {
        const uint64_t rou_inv_raw[4] = {0x0ed3e50a414e6dbaULL, 0xb22625f59115aba7ULL, 0x1bbe587180f34361ULL, 0x048127174daabc26ULL};  // fr.rs:93-98
        Fr ri = fp_to_mont<FrP>(fr_from_u64x4(rou_inv_raw));
        root = fp_inv<FrP>(ri);
        for (uint32_t i = k; i < 28; i++) root = fp_sqr<FrP>(root);
    }
    Fr n_fr = Fr::zero();
    n_fr.l[0] = (uint32_t)n; n_fr.l[1] = (uint32_t)((uint64_t)n >> 32);
    Fr n_inv = fp_inv<FrP>(fp_to_mont<FrP>(n_fr));
    Fr s_n = s;
    for (uint32_t i = 0; i < k; i++) s_n = fp_sqr<FrP>(s_n);  // s^(2^k)
    Fr mult = fp_mul<FrP>(fp_sub<FrP>(s_n, Fr::one()), n_inv);
    uint64_t root_limbs[4];
    for (int i = 0; i < 4; i++) root_limbs[i] = (uint64_t)root.l[2 * i] | ((uint64_t)root.l[2 * i + 1] << 32);
    CQB_TRY(fr_powers_run(root_limbs, n, d_pow));  // w^i
    unsigned grid = (unsigned)((n + 255) / 256);
    srs_diff_kernel<<<grid, 256, 0, st>>>((const uint4*)d_pow, n, s, (uint4*)d_aux);
    CQB_LAUNCHED();
    CQB_TRY(fr_batch_invert_run(d_aux, n));
    srs_lagrange_scalar_kernel<<<grid, 256, 0, st>>>((const uint4*)d_pow, (uint4*)d_aux, n, mult);
    CQB_LAUNCHED();
    CQB_TRY(g1_generator_mul_run(d_aux, n, d_g_lagrange));
    if (d_opening_at_0) {
        Fr root_inv = fp_inv<FrP>(root);
        uint64_t ri_limbs[4];
        for (int i = 0; i < 4; i++) ri_limbs[i] = (uint64_t)root_inv.l[2 * i] | ((uint64_t)root_inv.l[2 * i + 1] << 32);
        CQB_TRY(fr_powers_run(ri_limbs, n, d_aux2));  // w^-i  (commitment.rs:146-150 successors(..).batch_invert())
        // c = s^(N-1) / N : s^N / s ... computed as s_n * s^-1 * n_inv
        Fr c = fp_mul<FrP>(fp_mul<FrP>(s_n, fp_inv<FrP>(s)), n_inv);
        srs_opening_scalar_kernel<<<grid, 256, 0, st>>>((const uint4*)d_aux, (uint4*)d_aux2, n, c);
        CQB_LAUNCHED();
        CQB_TRY(g1_generator_mul_run(d_aux2, n, d_opening_at_0));
    }
    CQB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cqb

// ecntt.cu — SURVEY.md §8(a) row a18: g_to_lagrange (reference halo2_proofs/src/arithmetic.rs:277-301), the radix-2 FFT
// over G = G1 that ParamsKZG::downsize (poly/kzg/commitment.rs:482-490) uses to turn the monomial SRS g[0..n) into the
// Lagrange SRS: g_lagrange[i] = (1/n) sum_j omega^(-ij) g[j]. In the reference, group_scale is a 256-step double-and-add,
// so a butterfly costs ~3,500 Fq multiplications against 256 B of traffic: purely integer-bound, no tiling needed.
//   1. ec_load_bitrev : affine g[bitrev(i)] -> XYZZ work[i]                       (arithmetic.rs:186-191)
//   2. ec_stage (x log n): (a, b) <- (a + w b, a - w b), w = W[(i mod 2^s) << (L-1-s)] from the same resident twiddle
//      table the Fr NTT uses; w b by double-and-add over the canonical bits of w; twiddle one skipped (:213-219)
//   3. ec_scale_normalise : * n_inv (:286-290) and batch_normalize with one shared inversion per run (:292-298)
// Only the affine normal forms are canonical, so the result equals the reference's limb for limb.
#include "internal.h"

namespace cqb {

__device__ __forceinline__ Fq e_ld_fq(const uint4* p) {
    uint4 a = p[0], b = p[1];
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void e_st_fq(uint4* p, const Fq& v) {
    p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ G1Xyzz e_ld_pt(const uint4* p) {
    G1Xyzz r;
    r.x = e_ld_fq(p); r.y = e_ld_fq(p + 2); r.zz = e_ld_fq(p + 4); r.zzz = e_ld_fq(p + 6);
    return r;
}
__device__ __forceinline__ void e_st_pt(uint4* p, const G1Xyzz& v) {
    e_st_fq(p, v.x); e_st_fq(p + 2, v.y); e_st_fq(p + 4, v.zz); e_st_fq(p + 6, v.zzz);
}

// [k] P for a Montgomery-form scalar: the reference's group_scale (derive/curve.rs:914-935), MSB-first double-and-add
__device__ G1Xyzz g1_scalar_mul(const G1Xyzz& P, const Fr& k_mont) {
    Fr k = fp_from_mont<FrP>(k_mont);
    G1Xyzz acc = G1Xyzz::identity();
    bool started = false;
    for (int i = 7; i >= 0; i--)
        for (int b = 31; b >= 0; b--) {
            if (started) acc = g1_double(acc);
            if ((k.l[i] >> b) & 1u) { g1_add(acc, P); started = true; }
        }
    return acc;
}

__global__ void ec_load_bitrev_kernel(const uint4* __restrict__ g, uint4* __restrict__ work, uint32_t L) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << L)) return;
    uint32_t j = L ? (__brev(i) >> (32 - L)) : 0u;
    G1Affine a;
    a.x = e_ld_fq(g + (size_t)j * 4);
    a.y = e_ld_fq(g + (size_t)j * 4 + 2);
    e_st_pt(work + (size_t)i * 8, G1Xyzz::from_affine(a));
}

__global__ void __launch_bounds__(128) ec_stage_kernel(uint4* __restrict__ work, const uint4* __restrict__ tw, uint32_t L, uint32_t s) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (1u << (L - 1))) return;
    uint32_t i0 = ((b >> s) << (s + 1)) | (b & ((1u << s) - 1u));
    uint32_t i1 = i0 | (1u << s);
    uint32_t twi = (i0 & ((1u << s) - 1u)) << (L - 1 - s);
    G1Xyzz x = e_ld_pt(work + (size_t)i0 * 8), y = e_ld_pt(work + (size_t)i1 * 8);
    if (twi != 0) {
        Fr w;
        uint4 a = __ldg(tw + 2 * (size_t)twi), c = __ldg(tw + 2 * (size_t)twi + 1);
        w.l[0] = a.x; w.l[1] = a.y; w.l[2] = a.z; w.l[3] = a.w; w.l[4] = c.x; w.l[5] = c.y; w.l[6] = c.z; w.l[7] = c.w;
        y = g1_scalar_mul(y, w);
    }
    G1Xyzz u = x;
    g1_add(u, y);
    y.y = fp_neg<FqP>(y.y);
    g1_add(x, y);
    e_st_pt(work + (size_t)i0 * 8, u);
    e_st_pt(work + (size_t)i1 * 8, x);
}

constexpr int ECN_RUN = 16;
__global__ void __launch_bounds__(128) ec_scale_normalise_kernel(uint4* __restrict__ work, size_t n, Fr n_inv, uint4* __restrict__ prefix,
                                                                 uint4* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t p0 = t * ECN_RUN;
    if (p0 >= n) return;
    size_t cnt = (n - p0 < (size_t)ECN_RUN) ? (n - p0) : (size_t)ECN_RUN;
    Fq prod = Fq::one();
    for (size_t j = 0; j < cnt; j++) {
        G1Xyzz P = g1_scalar_mul(e_ld_pt(work + (p0 + j) * 8), n_inv);
        e_st_pt(work + (p0 + j) * 8, P);
        e_st_fq(prefix + (p0 + j) * 2, prod);
        if (!P.is_identity()) prod = fp_mul<FqP>(prod, fp_mul<FqP>(P.zz, P.zzz));
    }
    Fq inv = fp_inv<FqP>(prod);
    for (size_t j = cnt; j-- > 0;) {
        G1Xyzz P = e_ld_pt(work + (p0 + j) * 8);
        Fq ax = Fq::zero(), ay = Fq::zero();
        if (!P.is_identity()) {
            Fq zi = fp_mul<FqP>(inv, e_ld_fq(prefix + (p0 + j) * 2));
            inv = fp_mul<FqP>(inv, fp_mul<FqP>(P.zz, P.zzz));
            ax = fp_mul<FqP>(P.x, fp_mul<FqP>(zi, P.zzz));
            ay = fp_mul<FqP>(P.y, fp_mul<FqP>(zi, P.zz));
        }
        e_st_fq(out + (p0 + j) * 4, ax);
        e_st_fq(out + (p0 + j) * 4 + 2, ay);
    }
}

// ---- FK (Feist-Khovratovich) batch of all N KZG quotient commitments: the CQ table preprocessing -----------------------
// work[i] = XYZZ(i < n_in ? src_affine[bitrev-source] : identity): zero-padded bit-reversed load
__global__ void ec_load_bitrev_padded_kernel(const uint4* __restrict__ g, uint32_t n_in, uint4* __restrict__ work, uint32_t L) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << L)) return;
    uint32_t j = L ? (__brev(i) >> (32 - L)) : 0u;
    G1Xyzz P = G1Xyzz::identity();
    if (j < n_in) {
        G1Affine a;
        a.x = e_ld_fq(g + (size_t)j * 4);
        a.y = e_ld_fq(g + (size_t)j * 4 + 2);
        P = G1Xyzz::from_affine(a);
    }
    e_st_pt(work + (size_t)i * 8, P);
}
// dst[i] = src[bitrev(i)] (XYZZ)
__global__ void ec_bitrev_copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, uint32_t L) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << L)) return;
    uint32_t j = L ? (__brev(i) >> (32 - L)) : 0u;
    e_st_pt(dst + (size_t)i * 8, e_ld_pt(src + (size_t)j * 8));
}
// work[i] <- [sc[i]] work[i]
__global__ void __launch_bounds__(128) ec_pointwise_scale_kernel(uint4* __restrict__ work, const uint4* __restrict__ sc, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr k;
    uint4 a = sc[2 * i], c = sc[2 * i + 1];
    k.l[0] = a.x; k.l[1] = a.y; k.l[2] = a.z; k.l[3] = a.w; k.l[4] = c.x; k.l[5] = c.y; k.l[6] = c.z; k.l[7] = c.w;
    e_st_pt(work + i * 8, g1_scalar_mul(e_ld_pt(work + i * 8), k));
}
// tp[i] = t[d - i] * scale for i <= d = N-1, 0 for N <= i < 2N   (reversed, zero-padded coefficient vector)
__global__ void fk_reverse_pad_kernel(const uint4* __restrict__ t, uint32_t N, Fr scale, uint4* __restrict__ tp) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * N) return;
    Fr v = Fr::zero();
    if (i < N) {
        uint4 a = t[2 * (size_t)(N - 1 - i)], c = t[2 * (size_t)(N - 1 - i) + 1];
        v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w; v.l[4] = c.x; v.l[5] = c.y; v.l[6] = c.z; v.l[7] = c.w;
        v = fp_mul<FrP>(v, scale);
    }
    tp[2 * (size_t)i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    tp[2 * (size_t)i + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// H[i] (bit-reversed order, size N) = h_l with l = bitrev(i): h_l = conv[d-1-l] for l <= d-1, identity for l = N-1
__global__ void fk_gather_h_kernel(const uint4* __restrict__ conv, uint32_t N, uint32_t L, uint4* __restrict__ H) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t l = L ? (__brev(i) >> (32 - L)) : 0u;
    G1Xyzz P = G1Xyzz::identity();
    if (l + 1 < N) P = e_ld_pt(conv + (size_t)(N - 2 - l) * 8);
    e_st_pt(H + (size_t)i * 8, P);
}
// out[i] = affine([w^i / N] pi[i]); w^i from the half-size twiddle table (w^(i + N/2) = -w^i)
__global__ void __launch_bounds__(128) fk_finish_kernel(uint4* __restrict__ work, uint32_t N, const uint4* __restrict__ tw, Fr n_inv,
                                                         uint4* __restrict__ prefix, uint4* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t p0 = t * ECN_RUN;
    if (p0 >= N) return;
    size_t cnt = (N - p0 < (size_t)ECN_RUN) ? (N - p0) : (size_t)ECN_RUN;
    Fq prod = Fq::one();
    for (size_t j = 0; j < cnt; j++) {
        uint32_t i = (uint32_t)(p0 + j);
        Fr w = Fr::one();
        if (N > 1) {
            uint32_t half = N >> 1, ii = i & (half - 1);
            uint4 a = __ldg(tw + 2 * (size_t)ii), c = __ldg(tw + 2 * (size_t)ii + 1);
            w.l[0] = a.x; w.l[1] = a.y; w.l[2] = a.z; w.l[3] = a.w; w.l[4] = c.x; w.l[5] = c.y; w.l[6] = c.z; w.l[7] = c.w;
            if (i >= half) w = fp_neg<FrP>(w);
        }
        G1Xyzz P = g1_scalar_mul(e_ld_pt(work + (size_t)i * 8), fp_mul<FrP>(w, n_inv));
        e_st_pt(work + (size_t)i * 8, P);
        e_st_fq(prefix + (size_t)i * 2, prod);
        if (!P.is_identity()) prod = fp_mul<FqP>(prod, fp_mul<FqP>(P.zz, P.zzz));
    }
    Fq inv = fp_inv<FqP>(prod);
    for (size_t j = cnt; j-- > 0;) {
        size_t i = p0 + j;
        G1Xyzz P = e_ld_pt(work + i * 8);
        Fq ax = Fq::zero(), ay = Fq::zero();
        if (!P.is_identity()) {
            Fq zi = fp_mul<FqP>(inv, e_ld_fq(prefix + i * 2));
            inv = fp_mul<FqP>(inv, fp_mul<FqP>(P.zz, P.zzz));
            ax = fp_mul<FqP>(P.x, fp_mul<FqP>(zi, P.zzz));
            ay = fp_mul<FqP>(P.y, fp_mul<FqP>(zi, P.zz));
        }
        e_st_fq(out + i * 4, ax);
        e_st_fq(out + i * 4 + 2, ay);
    }
}

static Scratch g_ec_work, g_fk_a, g_fk_b, g_fk_sc;
void ecntt_release_all() { g_ec_work.release(); g_fk_a.release(); g_fk_b.release(); g_fk_sc.release(); }

static void fr_to_limbs64(const Fr& f, uint64_t out[4]) {
    for (int i = 0; i < 4; i++) out[i] = (uint64_t)f.l[2 * i] | ((uint64_t)f.l[2 * i + 1] << 32);
}
// butterfly stages of a size-2^L EC-NTT on `work` (already in bit-reversed order), root omega
static int ec_ntt_stages(uint4* work, uint32_t L, const Fr& omega) {
    if (L == 0) return 0;
    uint64_t w_limbs[4];
    fr_to_limbs64(omega, w_limbs);
    const void* tw = nullptr;
    CQB_TRY(ntt_get_twiddles(w_limbs, L, &tw));
    size_t n = (size_t)1 << L;
    for (uint32_t s = 0; s < L; s++) {
        ec_stage_kernel<<<(unsigned)((n / 2 + 127) / 128), 128, 0, ctx().stream>>>(work, (const uint4*)tw, L, s);
        CQB_LAUNCHED();
    }
    return 0;
}
static Fr root_of_unity_2pow(uint32_t L, bool inverse) {  // primitive 2^L-th root: ROOT_OF_UNITY^(2^(S-L)) (poly/domain.rs:54-61)
    const uint64_t rou_raw[4] = {0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL};      // fr.rs:77-82
    const uint64_t rou_inv_raw[4] = {0x0ed3e50a414e6dbaULL, 0xb22625f59115aba7ULL, 0x1bbe587180f34361ULL, 0x048127174daabc26ULL};  // fr.rs:93-98
    Fr w = fp_to_mont<FrP>(fr_from_u64x4(inverse ? rou_inv_raw : rou_raw));
    for (uint32_t i = L; i < 28; i++) w = fp_sqr<FrP>(w);
    return w;
}

// qs[i] = [ (T(X) - T(w^i)) / (X - w^i) * w^i / N ]_1 for all i < N = 2^log_n (reference plonk/static_lookup.rs:77-126, which
// does N kate_divisions + N MSMs of N-1 points and notes "TODO: THIS SHOULD BE DONE WITH FK METHOD" :107). FK: with
// t'_i = t_(N-1-i), h_l = (srs * t')[N-2-l] (one length-2N cyclic convolution through EC-NTTs) and pi = EC-NTT_N(h).
int cq_table_qs_run(const void* d_table_coeffs, uint32_t log_n, const void* d_srs_g1, void* d_qs_out) {
    if (log_n + 1 > 28) return fail(CQB_E_BAD_SIZE, "table of 2^%u rows needs a 2^%u transform (> Fr::S = 28)", log_n, log_n + 1);
    cudaStream_t st = ctx().stream;
    const uint32_t N = 1u << log_n, L2 = log_n + 1;
    const size_t N2 = (size_t)2 * N;
    CQB_TRY(g_fk_a.ensure(N2 * 128));
    CQB_TRY(g_fk_b.ensure(N2 * 128));
    CQB_TRY(g_fk_sc.ensure(N2 * 32 + (size_t)N * 32));
    uint4* A = g_fk_a.as<uint4>();
    uint4* B = g_fk_b.as<uint4>();
    uint4* sc = g_fk_sc.as<uint4>();
    uint4* prefix = sc + N2 * 2;
    const unsigned g2 = (unsigned)((N2 + 255) / 256), g1 = (unsigned)((N + 255) / 256);
    // A = EC-NTT_2N(srs padded)
    ec_load_bitrev_padded_kernel<<<g2, 256, 0, st>>>((const uint4*)d_srs_g1, N, A, L2);
    CQB_LAUNCHED();
    Fr w2 = root_of_unity_2pow(L2, false), w2_inv = root_of_unity_2pow(L2, true);
    CQB_TRY(ec_ntt_stages(A, L2, w2));
    // sc = NTT_2N(t' / 2N)   (the 1/2N of the inverse transform is folded into the scalars: field mul instead of point mul)
    Fr n2 = Fr::zero();
    n2.l[0] = (uint32_t)N2; n2.l[1] = (uint32_t)((uint64_t)N2 >> 32);
    Fr n2_inv = fp_inv<FrP>(fp_to_mont<FrP>(n2));
    fk_reverse_pad_kernel<<<g2, 256, 0, st>>>((const uint4*)d_table_coeffs, N, n2_inv, sc);
    CQB_LAUNCHED();
    uint64_t w2_limbs[4];
    fr_to_limbs64(w2, w2_limbs);
    NttFused none;
    CQB_TRY(ntt_run(sc, sc, L2, w2_limbs, none));
    // pointwise product, inverse EC-NTT
    ec_pointwise_scale_kernel<<<(unsigned)((N2 + 127) / 128), 128, 0, st>>>(A, sc, N2);
    CQB_LAUNCHED();
    ec_bitrev_copy_kernel<<<g2, 256, 0, st>>>(A, B, L2);
    CQB_LAUNCHED();
    CQB_TRY(ec_ntt_stages(B, L2, w2_inv));
    // h (bit-reversed) -> pi = EC-NTT_N(h)
    fk_gather_h_kernel<<<g1, 256, 0, st>>>(B, N, log_n, A);
    CQB_LAUNCHED();
    Fr w1 = root_of_unity_2pow(log_n, false);
    CQB_TRY(ec_ntt_stages(A, log_n, w1));
    // qs[i] = [w^i / N] pi[i], affine
    Fr n1 = Fr::zero();
    n1.l[0] = N;
    Fr n_inv = fp_inv<FrP>(fp_to_mont<FrP>(n1));
    const void* tw1 = nullptr;
    if (log_n >= 1) {
        uint64_t w1_limbs[4];
        fr_to_limbs64(w1, w1_limbs);
        CQB_TRY(ntt_get_twiddles(w1_limbs, log_n, &tw1));
    }
    size_t threads = ((size_t)N + ECN_RUN - 1) / ECN_RUN;
    fk_finish_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(A, N, (const uint4*)tw1, n_inv, prefix, (uint4*)d_qs_out);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

// d_g: n = 2^k affine points (monomial SRS prefix); d_out: n affine points (Lagrange SRS). May not alias.
int g_to_lagrange_run(const void* d_g, uint32_t k, void* d_out) {
    if (k > 28) return fail(CQB_E_BAD_SIZE, "k = %u exceeds Fr::S = 28", k);
    cudaStream_t st = ctx().stream;
    size_t n = (size_t)1 << k;
    // n_inv = TWO_INV^k ; omega_inv = ROOT_OF_UNITY_INV^(2^(S-k))   (arithmetic.rs:278-282)
    const uint64_t two_inv_raw[4] = {0xa1f0fac9f8000001ULL, 0x9419f4243cdcb848ULL, 0xdc2822db40c0ac2eULL, 0x183227397098d014ULL};      // fr.rs:85-90
    const uint64_t rou_inv_raw[4] = {0x0ed3e50a414e6dbaULL, 0xb22625f59115aba7ULL, 0x1bbe587180f34361ULL, 0x048127174daabc26ULL};      // fr.rs:93-98
    Fr two_inv = fp_to_mont<FrP>(fr_from_u64x4(two_inv_raw));
    Fr n_inv = Fr::one();
    for (uint32_t i = 0; i < k; i++) n_inv = fp_mul<FrP>(n_inv, two_inv);
    Fr omega_inv = fp_to_mont<FrP>(fr_from_u64x4(rou_inv_raw));
    for (uint32_t i = k; i < 28; i++) omega_inv = fp_sqr<FrP>(omega_inv);
    uint64_t w_limbs[4];
    for (int i = 0; i < 4; i++) w_limbs[i] = (uint64_t)omega_inv.l[2 * i] | ((uint64_t)omega_inv.l[2 * i + 1] << 32);
    CQB_TRY(g_ec_work.ensure(n * 128 + n * 32));
    uint4* work = g_ec_work.as<uint4>();
    uint4* prefix = work + n * 8;
    ec_load_bitrev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const uint4*)d_g, work, k);
    CQB_LAUNCHED();
    if (k >= 1) {
        const void* tw = nullptr;
        CQB_TRY(ntt_get_twiddles(w_limbs, k, &tw));
        for (uint32_t s = 0; s < k; s++) {
            ec_stage_kernel<<<(unsigned)((n / 2 + 127) / 128), 128, 0, st>>>(work, (const uint4*)tw, k, s);
            CQB_LAUNCHED();
        }
    }
    size_t threads = (n + ECN_RUN - 1) / ECN_RUN;
    ec_scale_normalise_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(work, n, n_inv, prefix, (uint4*)d_out);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cqb

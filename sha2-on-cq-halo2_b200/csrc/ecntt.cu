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

static Scratch g_ec_work;
void ecntt_release_all() { g_ec_work.release(); }

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

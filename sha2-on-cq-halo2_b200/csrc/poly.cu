// poly.cu — SURVEY.md §8(f) row 4: element-wise polynomial helpers that keep polynomials device-resident between the
// NTT / MSM calls of the prover:
//   eval_polynomial (reference halo2_proofs/src/arithmetic.rs:304-329): the reference splits the coefficient vector into
//       one chunk per thread, evaluates each by Horner and scales by point^start; same decomposition here with 64
//       coefficients per GPU thread and a block-wide sum.
//   kate_division   (reference arithmetic.rs:351-387): q_k = a_(k+1) + b q_(k+1), a serial recurrence in the reference;
//       here a chunked suffix scan: per-chunk Horner values, a recursive scan of the chunk carries with base b^C, then a
//       second pass that replays each chunk with its carry-in.
// Field arithmetic is exact, so any evaluation order gives the reference's limbs.
#include "internal.h"

namespace cqb {

__device__ __forceinline__ Fr p_ld(const uint4* p, size_t i) {
    uint4 a = p[2 * i], b = p[2 * i + 1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void p_st(uint4* p, size_t i, const Fr& v) {
    p[2 * i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[2 * i + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

constexpr int PCH = 64;  // coefficients per thread

// L[c] = sum_{k in chunk c} v[k] beta^(k - lo_c)   (Horner from the top of the chunk)
__global__ void __launch_bounds__(128) chunk_horner_kernel(const uint4* __restrict__ v, size_t m, Fr beta, uint4* __restrict__ L) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = c * PCH;
    if (lo >= m) return;
    size_t hi = (lo + PCH < m) ? lo + PCH : m;
    Fr acc = Fr::zero();
    for (size_t k = hi; k-- > lo;) acc = fp_add<FrP>(fp_mul<FrP>(acc, beta), p_ld(v, k));
    p_st(L, c, acc);
}
// S[k] = v[k] + beta S[k+1] inside chunk c, with carry-in Y[c+1] (0 for the top chunk). Y may be null (single chunk).
// v / S may alias (the recursive scan runs in place): no __restrict__ on them
__global__ void __launch_bounds__(128) chunk_replay_kernel(const uint4* v, size_t m, Fr beta, const uint4* Y,
                                                           size_t nchunks, uint4* S) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = c * PCH;
    if (lo >= m) return;
    size_t hi = (lo + PCH < m) ? lo + PCH : m;
    Fr acc = (Y != nullptr && c + 1 < nchunks) ? p_ld(Y, c + 1) : Fr::zero();
    for (size_t k = hi; k-- > lo;) {
        acc = fp_add<FrP>(fp_mul<FrP>(acc, beta), p_ld(v, k));
        p_st(S, k, acc);
    }
}
// partial[c] = L[c] * pw[c]   then block-summed by sum_kernel
__global__ void scale_by_powers_kernel(uint4* __restrict__ L, const uint4* __restrict__ pw, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    p_st(L, i, fp_mul<FrP>(p_ld(L, i), p_ld(pw, i)));
}
__global__ void __launch_bounds__(1024) fr_sum_kernel(const uint4* __restrict__ v, size_t n, uint4* __restrict__ out) {
    __shared__ uint4 sm[1024 * 2];
    Fr acc = Fr::zero();
    for (size_t i = threadIdx.x; i < n; i += 1024) acc = fp_add<FrP>(acc, p_ld(v, i));
    p_st(sm, threadIdx.x, acc);
    __syncthreads();
    for (int half = 512; half >= 1; half >>= 1) {
        if ((int)threadIdx.x < half) p_st(sm, threadIdx.x, fp_add<FrP>(p_ld(sm, threadIdx.x), p_ld(sm, threadIdx.x + half)));
        __syncthreads();
    }
    if (threadIdx.x == 0) p_st(out, 0, p_ld(sm, 0));
}

static Scratch g_poly_tmp;
void poly_release_all() { g_poly_tmp.release(); }

static Fr fr_pow_small(Fr b, unsigned e) {  // b^e for a small exponent (host)
    Fr r = Fr::one();
    while (e) {
        if (e & 1u) r = fp_mul<FrP>(r, b);
        b = fp_sqr<FrP>(b);
        e >>= 1;
    }
    return r;
}

// d_out (32 B device) = sum_i coeffs[i] point^i
int eval_polynomial_run(const void* d_coeffs, size_t n, const uint64_t point[4], void* d_out) {
    cudaStream_t st = ctx().stream;
    if (n == 0) {
        CQB_CUDA(cudaMemsetAsync(d_out, 0, 32, st));
        return 0;
    }
    Fr x = fr_from_u64x4(point);
    size_t nchunks = (n + PCH - 1) / PCH;
    CQB_TRY(g_poly_tmp.ensure(nchunks * 64 + 64));
    uint4* L = g_poly_tmp.as<uint4>();
    uint4* pw = L + nchunks * 2;
    chunk_horner_kernel<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>((const uint4*)d_coeffs, n, x, L);
    CQB_LAUNCHED();
    Fr xc = fr_pow_small(x, PCH);  // point^chunk: the reference's point.pow_vartime(&[start]) (:322), start = c * chunk
    uint64_t xc_limbs[4];
    for (int i = 0; i < 4; i++) xc_limbs[i] = (uint64_t)xc.l[2 * i] | ((uint64_t)xc.l[2 * i + 1] << 32);
    CQB_TRY(fr_powers_run(xc_limbs, nchunks, pw));
    scale_by_powers_kernel<<<(unsigned)((nchunks + 255) / 256), 256, 0, st>>>(L, pw, nchunks);
    CQB_LAUNCHED();
    fr_sum_kernel<<<1, 1024, 0, st>>>(L, nchunks, (uint4*)d_out);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

// S[k] = v[k] + beta S[k+1] for k < m (S[m] = 0); S may alias v. Recursive over chunk carries.
static int suffix_scan(const uint4* v, size_t m, Fr beta, uint4* S, size_t scratch_off) {
    cudaStream_t st = ctx().stream;
    size_t nchunks = (m + PCH - 1) / PCH;
    if (nchunks <= 1) {
        chunk_replay_kernel<<<1, 128, 0, st>>>(v, m, beta, nullptr, 1, S);
        CQB_LAUNCHED();
        return 0;
    }
    uint4* L = g_poly_tmp.as<uint4>() + scratch_off * 2;
    chunk_horner_kernel<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(v, m, beta, L);
    CQB_LAUNCHED();
    CQB_TRY(suffix_scan(L, nchunks, fr_pow_small(beta, PCH), L, scratch_off + nchunks));  // Y[c] = L[c] + beta^C Y[c+1], in place
    chunk_replay_kernel<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(v, m, beta, L, nchunks, S);
    CQB_LAUNCHED();
    return 0;
}

// q[0..n-1) = (a(X) - a(b)) / (X - b); d_q must not alias d_a
int kate_division_run(const void* d_a, size_t n, const uint64_t b[4], void* d_q) {
    if (n <= 1) return 0;
    size_t m = n - 1, total = 0;
    for (size_t c = (m + PCH - 1) / PCH; c > 1; c = (c + PCH - 1) / PCH) total += c;
    CQB_TRY(g_poly_tmp.ensure((total + 2) * 32 + 64));
    CQB_TRY(suffix_scan((const uint4*)d_a + 2, m, fr_from_u64x4(b), (uint4*)d_q, 0));  // v[k] = a[k+1]
    CQB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cqb

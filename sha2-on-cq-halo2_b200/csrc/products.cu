// products.cu — SURVEY.md §8(f) row 4: the grand products of the prover, device-resident.
//   fr_prefix_product : out[0] = init, out[i] = init * v[0] * ... * v[i-1]   (exclusive running product) — the reference
//       builds z by a serial loop (plonk/permutation/prover.rs:157-163; lookup/prover.rs has the same shape). Here a chunked
//       scan: per-chunk products, a recursive exclusive scan of the chunk products, then a replay of every chunk from its
//       carry-in. Field multiplication is associative and exact, so the limbs equal the serial loop's.
//   permutation_product : one column set of permutation::Argument::commit (permutation/prover.rs:82-166):
//       denominators prod_j (beta s_j + gamma + p_j) -> batch inversion -> numerators prod_j (delta^j omega^i beta + gamma + p_j)
//       (omega^i read from the resident NTT twiddle table: W[i] for i < n/2, -W[i - n/2] above) -> z by fr_prefix_product.
#include "internal.h"

namespace cqb {

__device__ __forceinline__ Fr q_ld(const uint4* p, size_t i) {
    uint4 a = p[2 * i], b = p[2 * i + 1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void q_st(uint4* p, size_t i, const Fr& v) {
    p[2 * i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[2 * i + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

constexpr int QCH = 32;  // elements per thread

__global__ void __launch_bounds__(128) chunk_product_kernel(const uint4* __restrict__ v, size_t m, uint4* __restrict__ L) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = c * QCH;
    if (lo >= m) return;
    size_t hi = (lo + QCH < m) ? lo + QCH : m;
    Fr acc = q_ld(v, lo);
    for (size_t k = lo + 1; k < hi; k++) acc = fp_mul<FrP>(acc, q_ld(v, k));
    q_st(L, c, acc);
}
// out[k] = carry * v[lo] * ... * v[k-1] for k in chunk c; carry = Y[c] (or init when Y is null: a single chunk).
// out may alias v (each element is read before its slot is written, by the same thread).
// in / out may alias (the recursive scans run in place): no __restrict__ on them
__global__ void __launch_bounds__(128) chunk_exclusive_replay_kernel(const uint4* v, size_t m, Fr init, const uint4* Y,
                                                                     uint4* out) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = c * QCH;
    if (lo >= m) return;
    size_t hi = (lo + QCH < m) ? lo + QCH : m;
    Fr acc = Y ? q_ld(Y, c) : init;
    for (size_t k = lo; k < hi; k++) {
        Fr x = q_ld(v, k);
        q_st(out, k, acc);
        acc = fp_mul<FrP>(acc, x);
    }
}

static Scratch g_prod_tmp, g_prod_mv;
void products_release_all() { g_prod_tmp.release(); g_prod_mv.release(); }

static int exclusive_product_scan(const uint4* v, size_t m, const Fr& init, uint4* out, size_t scratch_off) {
    cudaStream_t st = ctx().stream;
    size_t nchunks = (m + QCH - 1) / QCH;
    if (nchunks <= 1) {
        chunk_exclusive_replay_kernel<<<1, 128, 0, st>>>(v, m, init, nullptr, out);
        CQB_LAUNCHED();
        return 0;
    }
    uint4* L = g_prod_tmp.as<uint4>() + scratch_off * 2;
    chunk_product_kernel<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(v, m, L);
    CQB_LAUNCHED();
    CQB_TRY(exclusive_product_scan(L, nchunks, init, L, scratch_off + nchunks));  // L[c] <- init * prod of the chunks before c
    chunk_exclusive_replay_kernel<<<(unsigned)((nchunks + 127) / 128), 128, 0, st>>>(v, m, init, L, out);
    CQB_LAUNCHED();
    return 0;
}

// d_out[0] = init, d_out[i] = init * d_in[0] * ... * d_in[i-1], i < n. d_out may alias d_in.
int fr_prefix_product_run(const void* d_in, size_t n, const uint64_t init[4], void* d_out) {
    if (n == 0) return 0;
    size_t total = 0;
    for (size_t c = (n + QCH - 1) / QCH; c > 1; c = (c + QCH - 1) / QCH) total += c;
    CQB_TRY(g_prod_tmp.ensure((total + 2) * 32 + 64));
    CQB_TRY(exclusive_product_scan((const uint4*)d_in, n, fr_from_u64x4(init), (uint4*)d_out, 0));
    CQB_CUDA(cudaGetLastError());
    return 0;
}

constexpr int PERM_MAX_COLS = 16;
struct PermArgs {
    const uint4* col[PERM_MAX_COLS];
    const uint4* perm[PERM_MAX_COLS];
    uint32_t ncols;
    size_t n;
    Fr beta, gamma;
    Fr dw_beta[PERM_MAX_COLS];  // deltaomega_j * beta
    const uint4* tw;            // W[i] = omega^i, i < n/2
};
// mv[i] = prod_j (beta * perm_j[i] + gamma + col_j[i])      permutation/prover.rs:103-118
__global__ void __launch_bounds__(256) perm_denominator_kernel(const __grid_constant__ PermArgs a, uint4* __restrict__ mv) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    Fr acc = Fr::one();
    for (uint32_t j = 0; j < a.ncols; j++) {
        Fr t = fp_add<FrP>(fp_add<FrP>(fp_mul<FrP>(a.beta, q_ld(a.perm[j], i)), a.gamma), q_ld(a.col[j], i));
        acc = fp_mul<FrP>(acc, t);
    }
    q_st(mv, i, acc);
}
// mv[i] *= prod_j (deltaomega_j * omega^i * beta + gamma + col_j[i])      permutation/prover.rs:125-144
__global__ void __launch_bounds__(256) perm_numerator_kernel(const __grid_constant__ PermArgs a, uint4* __restrict__ mv) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const size_t half = a.n >> 1;
    Fr w;  // omega^i
    if (a.n == 1) w = Fr::one();
    else if (i < half) w = q_ld(a.tw, i);
    else w = fp_neg<FrP>(q_ld(a.tw, i - half));  // omega^(n/2) = -1
    Fr acc = q_ld(mv, i);
    for (uint32_t j = 0; j < a.ncols; j++) {
        Fr t = fp_add<FrP>(fp_add<FrP>(fp_mul<FrP>(a.dw_beta[j], w), a.gamma), q_ld(a.col[j], i));
        acc = fp_mul<FrP>(acc, t);
    }
    q_st(mv, i, acc);
}

// one column set; deltaomega_io: in = DELTA^(index of the set's first column), out = the value for the next set
int permutation_product_run(const void* const* d_columns, const void* const* d_perms, uint32_t ncols, uint32_t k, const uint64_t beta[4],
                            const uint64_t gamma[4], const uint64_t omega[4], const uint64_t delta[4], uint64_t deltaomega_io[4],
                            const uint64_t last_z[4], void* d_z) {
    if (ncols == 0 || ncols > PERM_MAX_COLS) return fail(CQB_E_BAD_ARG, "permutation product: 1..%d columns per set (got %u)", PERM_MAX_COLS, ncols);
    if (k > 28) return fail(CQB_E_BAD_SIZE, "log_n = %u exceeds Fr::S = 28", k);
    cudaStream_t st = ctx().stream;
    size_t n = (size_t)1 << k;
    PermArgs a;
    a.ncols = ncols;
    a.n = n;
    a.beta = fr_from_u64x4(beta);
    a.gamma = fr_from_u64x4(gamma);
    Fr dw = fr_from_u64x4(deltaomega_io), dl = fr_from_u64x4(delta);
    for (uint32_t j = 0; j < ncols; j++) {
        a.col[j] = (const uint4*)d_columns[j];
        a.perm[j] = (const uint4*)d_perms[j];
        a.dw_beta[j] = fp_mul<FrP>(dw, a.beta);
        dw = fp_mul<FrP>(dw, dl);
    }
    for (int i = 0; i < 4; i++) deltaomega_io[i] = (uint64_t)dw.l[2 * i] | ((uint64_t)dw.l[2 * i + 1] << 32);
    a.tw = nullptr;
    if (k >= 1) {
        const void* tw = nullptr;
        CQB_TRY(ntt_get_twiddles(omega, k, &tw));
        a.tw = (const uint4*)tw;
    }
    CQB_TRY(g_prod_mv.ensure(n * 32));
    uint4* mv = g_prod_mv.as<uint4>();
    unsigned grid = (unsigned)((n + 255) / 256);
    perm_denominator_kernel<<<grid, 256, 0, st>>>(a, mv);
    CQB_LAUNCHED();
    CQB_TRY(fr_batch_invert_run(mv, n));  // :121 modified_values.batch_invert()
    perm_numerator_kernel<<<grid, 256, 0, st>>>(a, mv);
    CQB_LAUNCHED();
    CQB_TRY(fr_prefix_product_run(mv, n, last_z, d_z));  // :157-163 (z has n entries: the product of all n fractions is not stored)
    CQB_CUDA(cudaGetLastError());
    return 0;
}


// ---- plookup grand product: lookup::prover::Permuted::commit_product (plonk/lookup/prover.rs:173-262) -------------------
// lookup_product[i] = (beta + a'_i)(gamma + s'_i)                                     (:208-216)
__global__ void __launch_bounds__(256) lookup_denominator_kernel(const uint4* __restrict__ pin, const uint4* __restrict__ ptab, size_t n, Fr beta,
                                                                 Fr gamma, uint4* __restrict__ lp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q_st(lp, i, fp_mul<FrP>(fp_add<FrP>(beta, q_ld(pin, i)), fp_add<FrP>(gamma, q_ld(ptab, i))));
}
// lookup_product[i] *= (a_i + beta)(s_i + gamma), a / s the theta-compressed input / table expressions   (:225-232)
__global__ void __launch_bounds__(256) lookup_numerator_kernel(const uint4* __restrict__ cin, const uint4* __restrict__ ctab, size_t n, Fr beta,
                                                               Fr gamma, uint4* __restrict__ lp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = fp_mul<FrP>(q_ld(lp, i), fp_add<FrP>(q_ld(cin, i), beta));
    q_st(lp, i, fp_mul<FrP>(v, fp_add<FrP>(q_ld(ctab, i), gamma)));
}
// d_z[0] = 1, d_z[i] = prod_{r < i} lookup_product[r], i < 2^k (:249-254; the caller overwrites the blinding rows, :259)
int lookup_product_run(const void* d_compressed_input, const void* d_compressed_table, const void* d_permuted_input, const void* d_permuted_table,
                       uint32_t k, const uint64_t beta[4], const uint64_t gamma[4], void* d_z) {
    if (k > 28) return fail(CQB_E_BAD_SIZE, "log_n = %u exceeds Fr::S = 28", k);
    cudaStream_t st = ctx().stream;
    size_t n = (size_t)1 << k;
    Fr b = fr_from_u64x4(beta), g = fr_from_u64x4(gamma);
    CQB_TRY(g_prod_mv.ensure(n * 32));
    uint4* lp = g_prod_mv.as<uint4>();
    unsigned grid = (unsigned)((n + 255) / 256);
    lookup_denominator_kernel<<<grid, 256, 0, st>>>((const uint4*)d_permuted_input, (const uint4*)d_permuted_table, n, b, g, lp);
    CQB_LAUNCHED();
    CQB_TRY(fr_batch_invert_run(lp, n));  // :220
    lookup_numerator_kernel<<<grid, 256, 0, st>>>((const uint4*)d_compressed_input, (const uint4*)d_compressed_table, n, b, g, lp);
    CQB_LAUNCHED();
    Fr one = Fr::one();
    uint64_t one_l[4];
    for (int i = 0; i < 4; i++) one_l[i] = (uint64_t)one.l[2 * i] | ((uint64_t)one.l[2 * i + 1] << 32);
    CQB_TRY(fr_prefix_product_run(lp, n, one_l, d_z));
    CQB_CUDA(cudaGetLastError());
    return 0;
}

// ---- element-wise pieces of the CQ prover (plonk/static_lookup/prover.rs), device-resident -------------------------------
constexpr int CQ_MAX_COLS = 16;
struct CompressArgs {
    const uint4* col[CQ_MAX_COLS];
    uint32_t ncols;
    size_t n;
    Fr theta;
    const uint32_t* idx;  // null: row i reads col[k][i]; else col[k][idx[i]]
};
// out[i] = fold_k (acc * theta + col_k[row]) starting from 0: compress_expressions (prover.rs:108-117, row = i) and the
// scalar half of compress_tables (prover.rs:224-229, row = idx[i])
__global__ void __launch_bounds__(256) fr_compress_kernel(const __grid_constant__ CompressArgs a, uint4* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    size_t row = a.idx ? (size_t)a.idx[i] : i;
    Fr acc = q_ld(a.col[0], row);  // 0 * theta + col_0
    for (uint32_t k = 1; k < a.ncols; k++) acc = fp_add<FrP>(fp_mul<FrP>(acc, a.theta), q_ld(a.col[k], row));
    q_st(out, i, acc);
}
// out[i] = in[i] + shift for i < usable, shift for i >= usable (the reference's bs before inversion, prover.rs:261-269)
__global__ void __launch_bounds__(256) fr_shift_kernel(const uint4* in, size_t n, size_t usable, Fr shift, uint4* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q_st(out, i, i < usable ? fp_add<FrP>(q_ld(in, i), shift) : shift);
}
__global__ void __launch_bounds__(256) fr_mul_kernel(const uint4* a, const uint4* b, size_t n, uint4* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q_st(out, i, fp_mul<FrP>(q_ld(a, i), q_ld(b, i)));
}

// acc[i] = acc[i] * a + x[i]: one Horner step over whole polynomials — h(X) = sum_i h_i(X) x^(n i) (vanishing/prover.rs:131-135) and
// the GWC batches sum_i v^i p_i(X) (poly/kzg/multiopen/gwc/prover.rs:62-77) are folds of it
__global__ void __launch_bounds__(256) fr_axpy_kernel(uint4* acc, Fr a, const uint4* x, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q_st(acc, i, fp_add<FrP>(fp_mul<FrP>(q_ld(acc, i), a), q_ld(x, i)));
}
int fr_axpy_run(void* d_acc, const uint64_t a[4], const void* d_x, size_t n) {
    if (n == 0) return 0;
    fr_axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>((uint4*)d_acc, fr_from_u64x4(a), (const uint4*)d_x, n);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

int fr_compress_run(const void* const* d_cols, uint32_t ncols, const uint32_t* d_idx, size_t n, const uint64_t theta[4], void* d_out) {
    if (ncols == 0 || ncols > CQ_MAX_COLS) return fail(CQB_E_BAD_ARG, "compress: 1..%d columns (got %u)", CQ_MAX_COLS, ncols);
    if (n == 0) return 0;
    CompressArgs a;
    for (uint32_t k = 0; k < ncols; k++) a.col[k] = (const uint4*)d_cols[k];
    a.ncols = ncols;
    a.n = n;
    a.theta = fr_from_u64x4(theta);
    a.idx = d_idx;
    fr_compress_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>(a, (uint4*)d_out);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}
// d_out[i] = (d_in[i] + shift)^-1 for i < usable, shift^-1 for usable <= i < n (zero stays zero, ff::BatchInvert)
int fr_inv_shifted_run(const void* d_in, size_t n, size_t usable, const uint64_t shift[4], void* d_out) {
    if (n == 0) return 0;
    fr_shift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>((const uint4*)d_in, n, usable, fr_from_u64x4(shift), (uint4*)d_out);
    CQB_LAUNCHED();
    CQB_TRY(fr_batch_invert_run(d_out, n));
    CQB_CUDA(cudaGetLastError());
    return 0;
}
int fr_mul_run(const void* d_a, const void* d_b, size_t n, void* d_out) {
    if (n == 0) return 0;
    fr_mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>((const uint4*)d_a, (const uint4*)d_b, n, (uint4*)d_out);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cqb

// internal.h — declarations shared between the library's translation units (not part of the public ABI).
#pragma once
#include <string.h>

#include <algorithm>

#include "context.h"
#include "ec.cuh"
#include "fp.cuh"

namespace cqb {

// ---- ntt.cu ----
constexpr int NTT_PRE_MAX_PUB = 16;
struct NttFused {
    size_t n_in = 0;  // number of valid input elements (0 => 2^log_n); the rest read as zero
    int pre_mode = 0; // 0 none | 1: x *= pre[j % 3] for j % 3 != 0 | 2: x *= pre[j & (pre_len - 1)]
    int pre_len = 0;
    Fr pre[NTT_PRE_MAX_PUB];
    int post_mode = 0; // 0 none | 1: x *= post[0] | 2: x *= post[i % 3]
    Fr post[3];
    // distributed four-step NTT (batched, out of place): see NttPassArgs in ntt.cu
    int in_map = 0, out_map = 0;
    unsigned in_s = 0, out_s = 0;
    unsigned long long in_A = 0, in_B = 0, out_A = 0, out_B = 0;
    const void* tw2 = nullptr;
    unsigned tw2_L = 0;
    unsigned long long tw2_row0 = 0;
    void* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // out_map == 2
    unsigned peer_rows_log = 0;
    unsigned long long peer_self_off = 0;
};
// batch > 1: `batch` independent transforms of 2^log_n contiguous elements each, stored back to back
int ntt_run(const void* d_src, void* d_dst, uint32_t log_n, const uint64_t omega[4], const NttFused& f, uint32_t batch = 1);
int fr_mul_omega_powers_run(void* d_a, size_t rows, size_t cols, size_t row0, const uint64_t omega[4], uint32_t log_n);
int fr_transpose_run(const void* d_in, void* d_out, size_t rows, size_t cols);
int fr_scale_table(void* d_a, size_t n, const void* d_tab, uint32_t len);
void ntt_release_all();
Fr fr_from_u64x4(const uint64_t* p);

// ---- msm.cu ----
// sum_i scalars[i] * bases[idx ? idx[i] : i]; d_out receives 64 B affine x||y followed by a uint32 identity flag
// windowed layout; d_bases = element 0 of the base set, point id = idx ? idx[i] : offset + i
// nparts: 0 = automatic (large MSMs are cut into point-range parts whose sort phase overlaps the previous part's bucket
// accumulation on a second stream); ready[p] (optional, with nparts > 0): event the sort of part p waits for (H2D of its scalars)
// feeder (optional, with nparts > 0, instead of ready): called on the enqueuing thread right before part p's sort is queued; starts
// the transfer of the scalars [lo, lo + cnt) and returns the event their sort must wait for
struct MsmFeeder {
    virtual int feed(int part, size_t lo, size_t cnt, cudaEvent_t* ready_out) = 0;
    virtual ~MsmFeeder() {}
};
int msm_run(const void* d_bases, size_t offset, const void* d_scalars, const uint32_t* d_idx, size_t n, void* d_out_xy_flag, int nparts = 0,
            const cudaEvent_t* ready = nullptr, MsmFeeder* feeder = nullptr);
// single bucket set over a precomputed table (row w = 2^(c w) * bases), table_n points per row
// batch > 1: d_scalars holds `batch` contiguous vectors of n scalars, d_out receives batch x 80 B; each MSM gets its own
// bucket set, all kernels run once for the whole batch (the latency-bound tails are shared)
int msm_run_precomputed(const void* d_table, size_t table_n, int c, size_t offset, const void* d_scalars, const uint32_t* d_idx,
                        size_t n, void* d_out_xy_flag, int batch = 1, int nparts = 0, const cudaEvent_t* ready = nullptr,
                        MsmFeeder* feeder = nullptr);
void msm_set_parts(int p);
void msm_set_accumulator(int mode);     // 0 auto | 1 XYZZ mixed additions | 2 batched affine streams | 3 affine tree
void msm_set_sort_mode(int m);           // 0 auto | 1 one-thread-per-scalar scatter | 2 partitioned sort whenever the shape allows
void msm_set_tree_levels(int levels);
int msm_last_tree_levels();
void msm_set_tree_slabs(int slabs);          // experiments: slabs per level of the affine tree (0 = automatic)
void msm_set_affine_segment(int seg_log);
void msm_set_affine_variant(int v);
// bounds[0..nparts]: the point ranges the parts of an MSM cover (small_first: host-pointer MSMs, see msm.cu)
void msm_part_bounds(size_t n, int nparts, bool small_first, size_t* bounds);
int msm_precompute_window_bits(size_t n);
int msm_windows_for(int c);
int msm_precompute_table(const void* d_bases, size_t n, int c, void* d_table);
int g1_sum_affine_run(const void* d_points, size_t n, void* d_out_xy_flag);
int g1_batch_normalize_run(const void* d_jacobian, size_t n, void* d_affine);
void msm_release_all();
void msm_release_scratch();                              // grow-only working buffers of the current device
size_t msm_working_set_bytes(size_t n, int nwin);         // what an MSM over a table of n x nwin points allocates besides its scalars
void msm_set_window_bits(int c);
void msm_set_profiling(bool on);
int msm_phase_ms(float* ms, int cap);

// ---- ecntt.cu ----
int ntt_get_twiddles(const uint64_t omega[4], uint32_t log_n, const void** d_table);  // defined in ntt.cu (cached table)
int g_to_lagrange_run(const void* d_g, uint32_t k, void* d_out);
int cq_table_qs_run(const void* d_table_coeffs, uint32_t log_n, const void* d_srs_g1, void* d_qs_out);
void ecntt_release_all();

// ---- evalh.cu ----
int graph_evaluate_run(const cqb_graph_t* g, const void* const* d_fixed, uint32_t n_fixed, const void* const* d_advice, uint32_t n_advice,
                       const void* const* d_instance, uint32_t n_instance, const uint64_t* challenges, uint32_t n_challenges,
                       const uint64_t* beta, const uint64_t* gamma, const uint64_t* theta, const uint64_t* y, void* d_values, uint64_t size,
                       int32_t rot_scale);
int cq_lookup_h_run(void* d_values, const void* d_b, const void* d_f, const void* d_l_active, const uint64_t* beta, const uint64_t* y, uint64_t size);
int permutation_h_run(void* d_values, uint64_t size, int32_t rot_scale, int32_t last_rotation, uint32_t chunk_len, const void* const* d_sets,
                      uint32_t nsets, const void* const* d_columns, const void* const* d_perm_cosets, uint32_t ncols, const void* d_l0,
                      const void* d_l_last, const void* d_l_active, const uint64_t* beta, const uint64_t* gamma, const uint64_t* y,
                      const uint64_t* extended_omega);
int lookup_h_run(void* d_values, const void* d_table_value, const void* d_product, const void* d_permuted_input, const void* d_permuted_table,
                 const void* d_l0, const void* d_l_last, const void* d_l_active, const uint64_t* beta, const uint64_t* gamma, const uint64_t* y,
                 uint64_t size, int32_t rot_scale);
void evalh_release_all();

// ---- poly.cu ----
int eval_polynomial_run(const void* d_coeffs, size_t n, const uint64_t point[4], void* d_out);
int kate_division_run(const void* d_a, size_t n, const uint64_t b[4], void* d_q);
void poly_release_all();

// ---- products.cu ----
int fr_prefix_product_run(const void* d_in, size_t n, const uint64_t init[4], void* d_out);
int permutation_product_run(const void* const* d_columns, const void* const* d_perms, uint32_t ncols, uint32_t k, const uint64_t beta[4],
                            const uint64_t gamma[4], const uint64_t omega[4], const uint64_t delta[4], uint64_t deltaomega_io[4],
                            const uint64_t last_z[4], void* d_z);
int fr_compress_run(const void* const* d_cols, uint32_t ncols, const uint32_t* d_idx, size_t n, const uint64_t theta[4], void* d_out);
int fr_inv_shifted_run(const void* d_in, size_t n, size_t usable, const uint64_t shift[4], void* d_out);
int fr_mul_run(const void* d_a, const void* d_b, size_t n, void* d_out);
int fr_axpy_run(void* d_acc, const uint64_t a[4], const void* d_x, size_t n);
int lookup_product_run(const void* d_compressed_input, const void* d_compressed_table, const void* d_permuted_input, const void* d_permuted_table,
                       uint32_t k, const uint64_t beta[4], const uint64_t gamma[4], void* d_z);
void products_release_all();

// ---- srs.cu ----
int fr_powers_run(const uint64_t base[4], size_t count, void* d_out);  // defined in ntt.cu
int fr_batch_invert_run(void* d_a, size_t n);
int g1_generator_mul_run(const void* d_scalars, size_t n, void* d_out);
int srs_setup_run(uint32_t k, const uint64_t s_limbs[4], void* d_g, void* d_g_lagrange, void* d_opening_at_0);
void srs_release_all();

// ---- g2.cu ----
int g2_mul_run(const void* d_bases_or_null, const void* d_scalars, size_t n, void* d_out_affine);  // affine 128 B each
int g2_msm_run(const void* d_bases, const void* d_scalars, size_t n, void* d_out_affine_flag);       // 128 B affine + uint32 identity flag
void g2_release_all();

// ---- gen.cu ----
int synth_scalars_run(uint64_t seed, size_t start, size_t n, void* d_out);
int synth_bases_run(uint64_t seed, size_t start, size_t n, void* d_out);
void gen_release_all();

}  // namespace cqb

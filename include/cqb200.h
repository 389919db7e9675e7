/* cqb200.h — C ABI of libcqb200.so: the B200-native replacement for the data-parallel hot path of
 * aleph-zero-foundation/sha2-on-cq-halo2 (BN254 G1 multi-scalar multiplication + Fr NTT behind every KZG commitment).
 *
 * The reference has NO FFI/plugin interface: the boundary is two generic Rust free functions and two
 * trait methods. Each entry point below names the reference interface it replaces (file:line in /root/reference); the
 * Rust shim that binds them is rust-shim/ (source only; no Rust toolchain in this image) and INTEGRATION.md.
 *
 * Data layout (identical to the reference's in-memory layout on a little-endian host, so the shim passes slices as-is):
 *   Fr / Fq   : 4 x uint64 little-endian limbs, Montgomery form (x * 2^256 mod p), always canonical (< p)
 *               (arithmetic/curves/src/bn256/fr.rs:22-25, derive/field.rs:302-308 to_raw_bytes)
 *   G1Affine  : x || y = 8 x uint64 (64 B); identity = all zero (derive/curve.rs:696-709, :667-685)
 * MSM results are returned as the AFFINE normal form + an identity flag (only the affine form is
 * canonical; the shim rebuilds G1 { x, y, z: one }).
 *
 * Conventions: every function returns 0 on success or a CQB_E_* code (the shim turns a non-zero code into the same
 * panic the reference would raise, e.g. assert_eq!(coeffs.len(), bases.len()) arithmetic.rs:133). There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with CQB_E_NO_DEVICE. Calls are serialised per
 * process by an internal mutex (re-entrant from multiple host threads; one process per GPU).
 */
#ifndef CQB200_H
#define CQB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CQB_API __attribute__((visibility("default")))
#else
#define CQB_API
#endif

#define CQB_OK 0
#define CQB_E_NO_DEVICE 1   /* no CUDA device / cqb_init not called */
#define CQB_E_CUDA 2        /* a CUDA runtime call failed (see cqb_last_error) */
#define CQB_E_BAD_ARG 3     /* NULL pointer, bad handle, offset+n beyond the registered bases ... */
#define CQB_E_LEN_MISMATCH 4 /* the reference's assert_eq!(coeffs.len(), bases.len())  arithmetic.rs:133 */
#define CQB_E_BAD_SIZE 5    /* the reference's assert_eq!(n, 1 << log_n) arithmetic.rs:184 / log_n > Fr::S = 28 */
#define CQB_E_OOM 6

typedef uint64_t cqb_bases_t; /* handle of a device-resident base set (an SRS: ParamsKZG.g / .g_lagrange, TableSRS.*) */

/* ---- lifetime -------------------------------------------------------------------------------------------------- */
CQB_API int cqb_init(int device);        /* bind this process to one GPU (one process per GPU); idempotent */
CQB_API void cqb_shutdown(void);
CQB_API const char* cqb_last_error(void);
CQB_API int cqb_device_count(void);
/* run all subsequent work on the caller's CUDA stream (cudaStream_t), e.g. torch.cuda.current_stream().cuda_stream;
 * NULL restores the library's own stream */
CQB_API int cqb_set_stream(void* cuda_stream);
/* ONE process, SEVERAL GPUs (SURVEY.md 8(b) `cqb_init(int n_devices)`; cqb_init(device) above keeps the one-process-per-GPU form
 * bench.py is launched in): devices 0 .. n_devices-1, device 0 primary (NTTs, polynomial helpers and plain base sets live there).
 * A base set registered with cqb_bases_register_sharded is split by contiguous point range over the devices — the
 * decomposition best_multiexp makes across threads, halo2_proofs/src/arithmetic.rs:137-153 — and every MSM over it runs on
 * all of them: one host thread per device, the partial points gathered on device 0 with peer copies and folded there
 * (arithmetic.rs:153). No torch, no NCCL on this path. */
CQB_API int cqb_init_multi(int n_devices);
CQB_API int cqb_active_devices(void);
CQB_API int cqb_sync(void);
CQB_API unsigned long long cqb_launch_count(void); /* kernels launched by this library so far (bench.py's gpu_launches) */

/* ---- SRS residency: replaces the host Vec<G1Affine> of ParamsKZG { g, g_lagrange } (poly/kzg/commitment.rs:31-39)
 *      and TableSRS { g1, g1_lagrange, g_lagrange_opening_at_0 } (:42-47) as MSM operands ------------------------- */
CQB_API int cqb_bases_register(const uint64_t* affine_xy, size_t n, cqb_bases_t* out);        /* host -> device copy */
CQB_API int cqb_bases_register_device(const void* d_affine_xy, size_t n, cqb_bases_t* out);  /* adopt device memory, no copy */
CQB_API int cqb_bases_register_sharded(const uint64_t* affine_xy, size_t n, cqb_bases_t* out); /* split over the devices of cqb_init_multi */
CQB_API int cqb_bases_free(cqb_bases_t h);
CQB_API size_t cqb_bases_len(cqb_bases_t h);
/* points [offset, offset + n) of a registered set back to the host (ParamsKZG::write_custom, poly/kzg/commitment.rs:366-380)
 * or into other device memory, stream-ordered (ParamsKZG::downsize, :482-490, truncates g before rebuilding g_lagrange) */
CQB_API int cqb_bases_download(cqb_bases_t h, size_t offset, size_t n, uint64_t* affine_xy_out);
CQB_API int cqb_bases_copy_dev(cqb_bases_t h, size_t offset, size_t n, void* d_affine_xy_out);
/* Build the resident table 2^(c w) P_i for every window w (nwin x n x 64 B of HBM: 12 GiB for a 2^24 SRS at c = 22) so
 * that all windows of an MSM over this set share ONE bucket set: fewer, wider windows and a single bucket reduction.
 * One-time cost per SRS (like computing g_lagrange in setup, poly/kzg/commitment.rs:234-262). window_bits = 0 picks c
 * from n. MSMs over >= 1/8 of the set then use the table automatically; results are identical either way. */
CQB_API int cqb_bases_precompute(cqb_bases_t h, int window_bits);
CQB_API int cqb_bases_drop_precomputed(cqb_bases_t h);
CQB_API int cqb_bases_precomputed_window_bits(cqb_bases_t h); /* c of the table, 0 if none: windows per point = 254/c + 1 */

/* ---- MSM: replaces best_multiexp (halo2_proofs/src/arithmetic.rs:132-159) as called by
 *      Params::commit_lagrange (poly/kzg/commitment.rs:496-504), ParamsProver::commit (:539-543),
 *      static_lookup/prover.rs:165,299,310, vanishing/prover.rs:58,101-105 ---------------------------------------- */
/* sum_i scalars[i] * bases[offset + i], scalars on the host */
CQB_API int cqb_msm_bn254_g1(cqb_bases_t b, size_t offset, const uint64_t* scalars, size_t n, uint64_t out_xy[8], int* is_inf);
/* same, scalars already in device memory (bench "value": inputs resident in HBM) */
CQB_API int cqb_msm_bn254_g1_dev(cqb_bases_t b, size_t offset, const void* d_scalars, size_t n, uint64_t out_xy[8], int* is_inf);
/* the same two calls with the result left on the device (80 bytes at d_out_xy_flag: affine x||y, then a uint32 identity flag),
 * queued on the library's stream and NOT waited for: a multi-process caller all-gathers its partial straight from there */
CQB_API int cqb_msm_bn254_g1_dev_to(cqb_bases_t b, size_t offset, const void* d_scalars, size_t n, void* d_out_xy_flag);
CQB_API int cqb_msm_bn254_g1_to(cqb_bases_t b, size_t offset, const uint64_t* scalars, size_t n, void* d_out_xy_flag);
/* sharded base set, scalars already resident: d_scalars[i] points, on shard i's device, at the scalars of that shard's point
 * range intersected with [offset, offset + n) (shard ranges: n / devices points each, the first n % devices one more) */
CQB_API int cqb_msm_bn254_g1_multi_dev(cqb_bases_t b, size_t offset, const void* const* d_scalars, size_t n, uint64_t out_xy[8], int* is_inf);
/* device memory on the device of slot `slot` (cqb_init_multi) for callers without a CUDA binding of their own */
CQB_API int cqb_dev_alloc_on(int slot, size_t bytes, void** d_out);
CQB_API int cqb_dev_free_on(int slot, void* d);
CQB_API int cqb_memcpy_h2d_on(int slot, void* d_dst, const void* h_src, size_t bytes);
CQB_API int cqb_synth_scalars_dev_on(int slot, uint64_t seed, size_t start, size_t n, void* d_out);
/* `batch` MSMs over the same base range in one pass: scalars = batch contiguous vectors of n scalars, out_xy = batch x 8
 * limbs, is_inf = batch flags. What the prover's commitment loops are (`advice.iter().map(|poly| params.commit_lagrange(poly))`
 * plonk/prover.rs:356-360; the h pieces vanishing/prover.rs:101-105): with a precomputed table every MSM gets its own bucket
 * set and all kernels run once for the whole batch, so the latency-bound tail is paid once. 1 <= batch <= 64. */
CQB_API int cqb_msm_bn254_g1_batch(cqb_bases_t b, size_t offset, const uint64_t* scalars, size_t n, int batch, uint64_t* out_xy, int* is_inf);
CQB_API int cqb_msm_bn254_g1_batch_dev(cqb_bases_t b, size_t offset, const void* d_scalars, size_t n, int batch, uint64_t* out_xy, int* is_inf);
/* one-shot: bases and scalars both on the host (exact best_multiexp(&[Fr], &[G1Affine]) shape) */
CQB_API int cqb_msm_bn254_g1_host(const uint64_t* affine_xy, const uint64_t* scalars, size_t n, uint64_t out_xy[8], int* is_inf);
/* cqb_msm_bn254_g1_host keeps large host base slices resident after their first use (key: pointer, length, fingerprint of the points;
 * an SRS slice is immutable for the life of its params): the commit loops of a prover stop re-sending 64 B per point per call. Budget in
 * bytes of device memory (default 1/8 of the device; 0 turns the cache off, as does CQB_HOST_BASES_CACHE=0 in the environment). */
CQB_API int cqb_set_host_bases_cache(long long budget_bytes);
/* sparse MSM sum_j scalars[j] * bases[idx[j]]: replaces the serial scalar-mul loops of the CQ prover for m(X), A(X),
 * Q_A(X), A_0(X) (plonk/static_lookup/prover.rs:167-170, 245-257) */
/* MSMKZG::eval (poly/kzg/msm.rs:65-70): projective bases (E::G1 = Jacobian x, y, z Montgomery limbs, 96 B each; z = 0 is the
 * identity) are normalised with Curve::batch_normalize (arithmetic/curves/src/derive/curve.rs:362-397) on the device, then multiplied */
CQB_API int cqb_g1_batch_normalize(const uint64_t* jacobian_xyz, size_t n, uint64_t* affine_xy_out);
CQB_API int cqb_msm_bn254_g1_jacobian(const uint64_t* jacobian_xyz, const uint64_t* scalars, size_t n, uint64_t out_xy[8], int* is_inf);
CQB_API int cqb_msm_bn254_g1_sparse(cqb_bases_t b, const uint32_t* idx, const uint64_t* scalars, size_t m, uint64_t out_xy[8],
                            int* is_inf);
/* sum of n affine points (host): the final fold of per-GPU partial results of a point-range-sharded MSM — the
 * multi-GPU analogue of results.iter().fold(identity, |a, b| a + b), arithmetic.rs:153 */
CQB_API int cqb_g1_sum_affine(const uint64_t* affine_xy, size_t n, uint64_t out_xy[8], int* is_inf);
CQB_API int cqb_g1_sum_affine_dev(const void* d_affine_xy, size_t n, uint64_t out_xy[8], int* is_inf); /* the partials already on the device (all-gather output) */

/* ---- NTT: replaces best_fft::<Fr> (halo2_proofs/src/arithmetic.rs:171-234) and its EvaluationDomain wrappers ---- */
/* in place, natural order in and out: a[k] <- sum_j a[j] omega^(jk); n = 1 << log_n (arithmetic.rs:184) */
CQB_API int cqb_ntt_bn254_fr(uint64_t* a, const uint64_t omega[4], uint32_t log_n);
CQB_API int cqb_ntt_bn254_fr_dev(void* d_a, const uint64_t omega[4], uint32_t log_n);
/* EvaluationDomain::ifft (poly/domain.rs:366-374): best_fft(omega_inv) then * divisor */
CQB_API int cqb_intt_bn254_fr(uint64_t* a, const uint64_t omega_inv[4], const uint64_t divisor[4], uint32_t log_n);
/* Building blocks of the distributed (multi-GPU) four-step NTT — ShardedNTT in the Python mirror; the exchange between the
 * steps is an all-to-all over NCCL / NVLink, outside this library:
 *   cqb_ntt_bn254_fr_batch_dev : `batch` (<= 65535) independent in-place transforms of 2^log_n contiguous elements each;
 *   cqb_fr_mul_omega_powers_dev: a[r][c] *= omega^((row0 + r) * c) over a rows x cols matrix, omega a 2^log_n-th root of unity
 *                                (the twiddle step between the column and the row transforms);
 *   cqb_fr_transpose_dev       : out[c][r] = in[r][c] for 32-byte elements (d_out must not alias d_in). */
CQB_API int cqb_ntt_bn254_fr_batch_dev(void* d_a, const uint64_t omega[4], uint32_t log_n, uint32_t batch);
/* The batched transform with the layout changes of the distributed NTT fused into its first gather and last store (out of
 * place): in_seg_log >= 0: member b's element idx is read from an all-to-all receive buffer laid out
 * [idx >> in_seg_log][b][idx & (2^in_seg_log - 1)] (source rank, member, segment); out_transposed != 0: results are stored
 * as [idx][b] (already in destination-rank order for the next all-to-all); tw_omega != NULL: result idx of member b is also
 * multiplied by tw_omega^((tw_row0 + b) * idx), tw_omega a 2^tw_log_n-th root of unity. */
CQB_API int cqb_ntt_bn254_fr_batch_map_dev(const void* d_src, void* d_dst, const uint64_t omega[4], uint32_t log_n, uint32_t batch,
                                           int in_seg_log, int out_transposed, const uint64_t tw_omega[4], uint32_t tw_log_n, size_t tw_row0,
                                           uint32_t in_batch_total /* 0 = batch; else members per segment of the source buffer */);
/* ... and with the exchange itself fused in: the last pass stores result idx of member b into the receive buffer of the rank
 * that owns it, peer_dst[idx >> (log_n - log2 n_peers)], at [self_rank][idx & mask][b] — peer memory over NVLink (pointers
 * from cqb_ipc_open; peer_dst[self_rank] is the caller's own buffer). d_scratch (batch * 2^log_n elements, local) holds the
 * intermediate passes. The caller synchronises the ranks before the buffers are read (any stream-ordered collective). */
CQB_API int cqb_ntt_bn254_fr_batch_p2p_dev(const void* d_src, void* d_scratch, void* const* peer_dst, uint32_t n_peers, uint32_t self_rank,
                                           const uint64_t omega[4], uint32_t log_n, uint32_t batch, int in_seg_log,
                                           const uint64_t tw_omega[4], uint32_t tw_log_n, size_t tw_row0);
/* CUDA IPC plumbing (one process per GPU): export a cqb_dev_alloc'ed buffer as a 64-byte handle, open a peer's handle, close it */
CQB_API int cqb_ipc_export(const void* d_ptr, unsigned char handle_out[64]);
CQB_API int cqb_ipc_open(const unsigned char handle[64], void** d_out);
CQB_API int cqb_ipc_close(void* d_ptr);
CQB_API int cqb_fr_mul_omega_powers_dev(void* d_a, size_t rows, size_t cols, size_t row0, const uint64_t omega[4], uint32_t log_n);
CQB_API int cqb_fr_transpose_dev(const void* d_in, void* d_out, size_t rows, size_t cols);
CQB_API int cqb_intt_bn254_fr_dev(void* d_a, const uint64_t omega_inv[4], const uint64_t divisor[4], uint32_t log_n);
/* EvaluationDomain::coeff_to_extended (poly/domain.rs:252-266): a[i] *= {1, g_coset, g_coset_inv}[i % 3]
 * (distribute_powers_zeta :347-363), zero-pad n -> 2^ext_log_n, best_fft(extended_omega). out has 2^ext_log_n elements. */
CQB_API int cqb_coset_ntt_bn254_fr(const uint64_t* coeffs, size_t n, uint64_t* out, const uint64_t ext_omega[4],
                           uint32_t ext_log_n, const uint64_t g_coset[4], const uint64_t g_coset_inv[4]);
CQB_API int cqb_coset_ntt_bn254_fr_dev(const void* d_coeffs, size_t n, void* d_out, const uint64_t ext_omega[4],
                               uint32_t ext_log_n, const uint64_t g_coset[4], const uint64_t g_coset_inv[4]);
/* EvaluationDomain::divide_by_vanishing_poly (poly/domain.rs:319-338, skipped when t_evaluations == NULL) followed by
 * extended_to_coeff (:293-315): a[i] *= t_evaluations[i % t_len]; ifft(extended_omega_inv, extended_ifft_divisor);
 * a[i] *= {1, g_coset_inv, g_coset}[i % 3]. In place on 2^ext_log_n elements; the caller truncates to
 * n * quotient_poly_degree as the reference does (:311-312). */
CQB_API int cqb_coset_intt_bn254_fr(uint64_t* a, uint32_t ext_log_n, const uint64_t ext_omega_inv[4],
                            const uint64_t ext_divisor[4], const uint64_t g_coset[4], const uint64_t g_coset_inv[4],
                            const uint64_t* t_evaluations, uint32_t t_len);
CQB_API int cqb_coset_intt_bn254_fr_dev(void* d_a, uint32_t ext_log_n, const uint64_t ext_omega_inv[4],
                                const uint64_t ext_divisor[4], const uint64_t g_coset[4], const uint64_t g_coset_inv[4],
                                const uint64_t* t_evaluations, uint32_t t_len);

/* ---- synthetic inputs generated on the device; same definition as oracle_synth_* ------------- */
/* scalars[i] = Fr::from_u512(splitmix64 stream(seed, start + i))  (mirrors Fr::random, bn256/fr.rs:159-170) */
CQB_API int cqb_synth_scalars_dev(uint64_t seed, size_t start, size_t n, void* d_out);
/* bases[i] = [s0 + (start + i) d] G, (s0, d) = first two scalars of stream `seed`; affine, normalised on the device */
CQB_API int cqb_synth_bases_dev(uint64_t seed, size_t start, size_t n, void* d_out);

/* ---- SRS generation + element-wise helpers ------------------------------------------
 * cqb_srs_setup_dev replaces the scalar-multiplication loops of ParamsKZG::setup_from_toxic_waste / setup
 * (poly/kzg/commitment.rs:209-276, 280-348; identical G1 formulas in TableSRS::setup_from_toxic_waste :73-141):
 * d_g[i] = [s^i]G, d_g_lagrange[i] = [(s^n - 1)/n * w^i/(s - w^i)]G for n = 2^k, affine, 64 B each, written to device
 * memory (register them with cqb_bases_register_device — the SRS never has to exist on the host). */
CQB_API int cqb_srs_setup_dev(uint32_t k, const uint64_t s[4], void* d_g, void* d_g_lagrange);
/* G1 parts of TableSRS::setup_from_toxic_waste (poly/kzg/commitment.rs:73-178) for a table SRS of 2^log_len powers:
 * g1, g1_lagrange and g_lagrange_opening_at_0[i] = [(L_i(x) - L_i(0))/x]_1 (:143-170), what the CQ prover's m / A / A_0
 * commitments index into (plonk/static_lookup/prover.rs:167-170, 245-257) */
CQB_API int cqb_table_srs_setup_dev(uint32_t log_len, const uint64_t s[4], void* d_g1, void* d_g1_lagrange, void* d_opening_at_0);
/* g_to_lagrange (halo2_proofs/src/arithmetic.rs:277-301): radix-2 FFT over G1 of the first 2^k monomial SRS points, scaled
 * by 1/n and normalised -> the Lagrange SRS; what ParamsKZG::downsize needs (poly/kzg/commitment.rs:482-490).
 * d_g and d_out: 2^k affine points each, must not alias. */
CQB_API int cqb_g_to_lagrange_dev(const void* d_g, uint32_t k, void* d_out);
/* CQ table preprocessing: all N = 2^log_n cached quotient commitments of StaticTableValues::new
 * (plonk/static_lookup.rs:77-126): qs[i] = [ (T(X) - T(w^i))/(X - w^i) * w^i/N ]_1, T given by its N coefficients
 * (= ifft of the table values, :99-105), d_srs_g1 = the first N powers [x^j]_1. The reference runs N kate_divisions and N
 * MSMs of N-1 points (O(N^2), "TODO: THIS SHOULD BE DONE WITH FK METHOD" :107); this is the FK algorithm: three G1
 * EC-NTTs (sizes 2N, 2N, N) + 3N scalar multiplications, O(N log N). Output: N affine points (device). */
CQB_API int cqb_cq_table_qs_dev(const void* d_table_coeffs, uint32_t log_n, const void* d_srs_g1, void* d_qs_out);
/* out[i] = [scalars[i]] G (fixed-base batch multiplication by the bn256 generator (1,2)), affine */
CQB_API int cqb_g1_generator_mul_dev(const void* d_scalars, size_t n, void* d_out);
/* eval_polynomial (halo2_proofs/src/arithmetic.rs:304-329): out = sum_i coeffs[i] point^i, coefficients device-resident */
CQB_API int cqb_eval_polynomial_dev(const void* d_coeffs, size_t n, const uint64_t point[4], uint64_t out[4]);
/* count polynomials of n coefficients each, evaluated at points[i] (count x 4 limbs), one read-back for all (plonk/prover.rs:629-719) */
CQB_API int cqb_eval_polynomials_dev(const void* const* d_coeffs, size_t n, const uint64_t* points, uint32_t count, uint64_t* out);
/* kate_division (arithmetic.rs:351-387): d_q[0..n-1) = (a(X) - a(b)) / (X - b); as used by the multiopen provers
 * (poly/kzg/multiopen/gwc/prover.rs:80-86) and the CQ table preprocessing; d_q must not alias d_a */
CQB_API int cqb_kate_division_dev(const void* d_a, size_t n, const uint64_t b[4], void* d_q);
/* plookup (the original halo2 lookup argument; the CQ circuits of this repository use static lookups instead, but
 * create_proof runs both):
 *   cqb_lookup_product_dev: lookup::prover::Permuted::commit_product (plonk/lookup/prover.rs:173-262) — the grand product z of
 *     (a_i + beta)(s_i + gamma) / ((a'_i + beta)(s'_i + gamma)) over the theta-compressed input / table expressions and their
 *     permuted versions (2^k Lagrange values each, device-resident): d_z[0] = 1, d_z[i] = product of the first i fractions.
 *     The caller overwrites d_z[n - blinding_factors..] with its random blinding rows (:259).
 *   cqb_lookup_h_dev: the five plookup constraints of evaluate_h (plonk/evaluation.rs:458-531) folded into d_values with y;
 *     d_table_value = the lookup's GraphEvaluator output (compressed input + beta)(compressed table + gamma) per extended row,
 *     e.g. from cqb_graph_evaluate_dev on a zeroed vector. */
CQB_API int cqb_lookup_product_dev(const void* d_compressed_input, const void* d_compressed_table, const void* d_permuted_input,
                                   const void* d_permuted_table, uint32_t k, const uint64_t beta[4], const uint64_t gamma[4], void* d_z);
CQB_API int cqb_lookup_h_dev(void* d_values, const void* d_table_value, const void* d_product_coset, const void* d_permuted_input_coset,
                             const void* d_permuted_table_coset, const void* d_l0, const void* d_l_last, const void* d_l_active_row,
                             const uint64_t beta[4], const uint64_t gamma[4], const uint64_t y[4], uint64_t size, int32_t rot_scale);
/* Element-wise pieces of the CQ prover (plonk/static_lookup/prover.rs) on device-resident vectors:
 *   cqb_fr_compress_dev   : d_out[i] = fold_k (acc * theta + cols[k][row]) from acc = 0, row = d_idx ? d_idx[i] : i — the
 *                           theta-compression of the lookup's input expressions (:108-117, d_idx NULL) and of the table values
 *                           over the support of m (:224-229, d_idx = support, device pointer); ncols <= 16.
 *   cqb_fr_inv_shifted_dev: d_out[i] = (d_in[i] + shift)^-1 for i < usable, shift^-1 above — B's evaluations 1/(f_i + beta)
 *                           with 1/beta on the blinding rows (:261-269); with usable = n, the 1/(t_i + beta) of A (:243).
 *   cqb_fr_mul_dev        : element-wise product (a_i = multiplicity_i * 1/(t_i + beta), :243).
 *   cqb_msm_bn254_g1_sparse_dev: cqb_msm_bn254_g1_sparse with device-resident indices and scalars (indices are not range-checked). */
CQB_API int cqb_fr_compress_dev(const void* const* d_cols, uint32_t ncols, const uint32_t* d_idx, size_t n, const uint64_t theta[4], void* d_out);
CQB_API int cqb_fr_inv_shifted_dev(const void* d_in, size_t n, size_t usable, const uint64_t shift[4], void* d_out);
CQB_API int cqb_fr_mul_dev(const void* d_a, const void* d_b, size_t n, void* d_out);
/* d_acc[i] = d_acc[i] * a + d_x[i]: one Horner step over whole polynomials. h(X) = sum_i h_i(X) x^(n i) (plonk/vanishing/prover.rs:131-135)
 * and the GWC batches sum_i v^i p_i(X) (poly/kzg/multiopen/gwc/prover.rs:62-77) are folds of it. */
CQB_API int cqb_fr_axpy_dev(void* d_acc, const uint64_t a[4], const void* d_x, size_t n);
CQB_API int cqb_msm_bn254_g1_sparse_dev(cqb_bases_t b, const uint32_t* d_idx, const void* d_scalars, size_t m, uint64_t out_xy[8], int* is_inf);
/* Exclusive running product, the serial z-loop of the grand-product arguments (plonk/permutation/prover.rs:157-163):
 * d_out[0] = init, d_out[i] = init * d_in[0] * ... * d_in[i-1] for i < n. d_out may alias d_in. */
CQB_API int cqb_fr_prefix_product_dev(const void* d_in, size_t n, const uint64_t init[4], void* d_out);
/* One column set of permutation::Argument::commit (plonk/permutation/prover.rs:82-166): the grand-product vector z (2^k
 * Lagrange values) of `ncols` (<= 16 = chunk_len) columns d_columns[j] against their permutation polynomials d_perms[j]
 * (pkey.permutations, Lagrange values), all 2^k Fr on the device:
 *   z[0] = last_z, z[i+1] = z[i] * prod_j (col_j[i] + delta^j' omega^i beta + gamma) / (col_j[i] + beta perm_j[i] + gamma).
 * deltaomega_io: in = DELTA^(index of the set's first column), out = the value for the next set (:144); delta = Fr::DELTA;
 * omega = the domain generator. The blinding rows z[n - blinding_factors..] are overwritten by the caller from its rng
 * (:152-155), and last_z of the next set is z[n - (blinding_factors + 1)] (:157). */
CQB_API int cqb_permutation_product_dev(const void* const* d_columns, const void* const* d_perms, uint32_t ncols, uint32_t k,
                                        const uint64_t beta[4], const uint64_t gamma[4], const uint64_t omega[4], const uint64_t delta[4],
                                        uint64_t deltaomega_io[4], const uint64_t last_z[4], void* d_z);
/* in place a[i] <- a[i] * factor (the parallelize()d scaling loops, e.g. poly/domain.rs:369-373) */
CQB_API int cqb_fr_scale_dev(void* d_a, size_t n, const uint64_t factor[4]);
/* in place a[i] <- 1/a[i], zeros stay zero: ff::BatchInvert as used at poly/domain.rs:118-125, static_lookup/prover.rs:261-269 */
CQB_API int cqb_fr_batch_invert_dev(void* d_a, size_t n);
/* out[i] = base^i, i < n (the serial scans at arithmetic.rs:194-200, commitment.rs:153-156) */
CQB_API int cqb_fr_powers_dev(const uint64_t base[4], size_t n, void* d_out);

/* ---- evaluate_h on the device (halo2_proofs/src/plonk/evaluation.rs:285-551) ---------------------
 * A GraphEvaluator (evaluation.rs:197-207) is passed serialised (host memory):
 *   constants  : n_constants Fr (Montgomery limbs); rotations: n_rotations int32 (Rotation.0)
 *   code       : the calculations in order, as 32-bit words. A ValueSource (:41-65) is 2 words
 *                { kind | rotation_index << 8 , index }, kind = 0 Constant, 1 Intermediate, 2 Fixed, 3 Advice, 4 Instance,
 *                5 Challenge, 6 Beta, 7 Gamma, 8 Theta, 9 Y, 10 PreviousValue (index = constant / intermediate / column /
 *                challenge index). A Calculation (:114-132) is { op, target, operands... }, op = 0 Add(a,b), 1 Sub(a,b),
 *                2 Mul(a,b), 3 Square(a), 4 Double(a), 5 Negate(a), 6 Horner(start, factor, nparts, parts...), 7 Store(a).
 * Limits: <= 32 rotations, <= 512 intermediates. */
typedef struct {
    const uint64_t* constants; uint32_t n_constants;
    const int32_t* rotations;  uint32_t n_rotations;
    const uint32_t* code;      uint32_t code_words;
    uint32_t n_calculations;   uint32_t num_intermediates;
} cqb_graph_t;
/* GraphEvaluator::evaluate (:718-775) for every row idx < size of the extended domain, in place on d_values
 * (values[idx] is the PreviousValue and receives the result; custom gates :348-374). d_fixed / d_advice / d_instance: HOST
 * arrays of DEVICE pointers to the coset evaluations (size Fr each). */
CQB_API int cqb_graph_evaluate_dev(const cqb_graph_t* graph, const void* const* d_fixed, uint32_t n_fixed, const void* const* d_advice,
                                   uint32_t n_advice, const void* const* d_instance, uint32_t n_instance, const uint64_t* challenges,
                                   uint32_t n_challenges, const uint64_t beta[4], const uint64_t gamma[4], const uint64_t theta[4],
                                   const uint64_t y[4], void* d_values, uint64_t size, int32_t rot_scale);
/* the static-lookup (CQ) term (:533-548): values = values * y + (b_coset * (f_coset * l_active_row + beta) - 1) */
CQB_API int cqb_cq_lookup_h_dev(void* d_values, const void* d_b_coset, const void* d_f_coset, const void* d_l_active_row,
                                const uint64_t beta[4], const uint64_t y[4], uint64_t size);
/* the permutation argument terms (:376-452); d_sets: nsets product cosets, d_columns / d_perm_cosets: ncols column value /
 * permutation cosets grouped chunk_len per set (HOST arrays of DEVICE pointers) */
CQB_API int cqb_permutation_h_dev(void* d_values, uint64_t size, int32_t rot_scale, int32_t last_rotation, uint32_t chunk_len,
                                  const void* const* d_sets, uint32_t nsets, const void* const* d_columns,
                                  const void* const* d_perm_cosets, uint32_t ncols, const void* d_l0, const void* d_l_last,
                                  const void* d_l_active_row, const uint64_t beta[4], const uint64_t gamma[4], const uint64_t y[4],
                                  const uint64_t extended_omega[4]);

/* ---- G2 (keygen-time, primary device). G2Affine = 128 bytes x.c0 || x.c1 || y.c0 || y.c1 (Fq Montgomery limbs, the reference's raw
 *      layout, arithmetic/curves/src/bn256/fq2.rs + derive/curve.rs SerdeObject); identity = zeros.
 *   cqb_g2_powers     : out[i] = [s^i] G2, i < count — ParamsKZG's s_g2 = out[1] (poly/kzg/commitment.rs:265-266, 337-338) and the table
 *                       SRS's G2 powers (:94-104, 114-141).
 *   cqb_msm_bn254_g2  : best_multiexp::<G2Affine> — the CQ table commitment t (plonk/static_lookup.rs:146); zv = [s^N]G2 - G2 (:137)
 *                       is the same call with scalars (1, r - 1).
 *   cqb_g2_generator_mul_dev: out[i] = [scalars[i]] G2 with device-resident scalars and output. */
CQB_API int cqb_g2_powers(const uint64_t s[4], size_t count, uint64_t* g2_affine_out);
CQB_API int cqb_msm_bn254_g2(const uint64_t* g2_affine, const uint64_t* scalars, size_t n, uint64_t out_xy[16], int* is_inf);
CQB_API int cqb_g2_generator_mul_dev(const void* d_scalars, size_t n, void* d_out_affine);

/* ---- plain device memory helpers so non-CUDA hosts (ctypes, the Rust shim) need no CUDA binding of their own ---- */
CQB_API int cqb_dev_alloc(size_t bytes, void** d_out);
CQB_API int cqb_dev_free(void* d);
CQB_API int cqb_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes);
CQB_API int cqb_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes);
CQB_API int cqb_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes); /* stream-ordered, asynchronous */
CQB_API int cqb_host_alloc_pinned(size_t bytes, void** h_out);
CQB_API int cqb_host_free_pinned(void* h);

/* measurement hooks: when profiling is on, every MSM records CUDA events around its kernels on the streams they run on;
 * cqb_msm_phase_ms returns how many of ms[0..7] = {count, scan, scatter, accumulate, merge, reduce, window_sum, final} it filled
 * for the most recent MSM (device time, milliseconds, summed over the parts of a pipelined MSM — sort phases of part p+1
 * overlap the accumulation of part p, so the phases can add up to more than the MSM's wall time) */
CQB_API int cqb_msm_set_profiling(int on);
CQB_API int cqb_msm_phase_ms(float* ms, int cap);

/* tuning knobs for experiments (0 = automatic): MSM window bits */
CQB_API int cqb_msm_set_window_bits(int c);
/* number of point-range parts a large device-resident MSM is cut into (the sort phase of part p+1 runs on a second stream
 * under the bucket accumulation of part p); 0 = automatic, 1 = no pipelining, at most 8. Results do not depend on it. */
CQB_API int cqb_msm_set_parts(int parts);
/* bucket accumulation variant (the CPU form of 2 and 3 is the reference's batch_add, arithmetic/curves/src/derive/curve.rs:4-141):
 *   0 = automatic: the affine tree (3) for MSMs over a precomputed table with >= 40 entries per bucket, its depth chosen from the
 *       bucket length; XYZZ otherwise
 *   1 = XYZZ mixed additions (1,232 MAD32 each)
 *   2 = batched affine additions as per-thread streams with one safegcd inversion per step: measured SLOWER than XYZZ on B200
 *       (2^24: 44-56 ms against 32 ms); kept selectable and parity-tested (tests/test_gpu_affine_acc.py)
 *   3 = the affine tree with the depth of cqb_msm_set_tree_levels: every bucket's run of the sorted list is padded to a multiple of
 *       2^levels entries, so that each level is ONE batch of independent pair additions (788 MAD32 each) over the whole list —
 *       forward pass (denominators, running products), one lane-parallel inversion pass, backward pass — and only 1 / 2^levels of
 *       the additions stay XYZZ. Measured 2^24: accumulation 28.1 ms against 32.0 ms (tests/test_gpu_affine_tree.py).
 * affine_seg_log: variant 2 only, entries per stream = 2^affine_seg_log (0 = automatic). Results do not depend on any of these. */
CQB_API int cqb_msm_set_accumulator(int mode, int affine_seg_log);
/* depth of the affine tree when variant 3 is forced: 1..6, default 4 */
CQB_API int cqb_msm_set_tree_levels(int levels);
/* sorting of the digits into bucket order: 0 = automatic = 1 = the one-thread-per-scalar scatter; 2 = the partitioned sort (tile-local sort by
 * the top 9 bits of the bucket id in shared memory, then per-bin count and placement) whenever the shape allows (window bits 11..22, one
 * bucket set). Measured slower on B200 (2^24: 6.2 ms against 4.5 ms, profiles/r02_partitioned_sort.md); selectable and parity-tested.
 * Results do not depend on it. */
CQB_API int cqb_msm_set_sort_mode(int mode);
/* tree depth the most recent MSM's accumulation ran with (its largest part); 0 = XYZZ mixed additions only */
CQB_API int cqb_msm_last_tree_levels(void);

#ifdef __cplusplus
}
#endif
#endif /* CQB200_H */

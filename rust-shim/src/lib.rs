//! Rust side of the drop-in boundary — a SOURCE SKETCH: never compiled (no Rust toolchain in this repository's image, see Cargo.toml);
//! `mod generic` below holds `unimplemented!()` placeholders where the reference's own generic bodies stay. What is tested is the C ABI
//! these declarations bind (tests/cpp/*.cpp, tests/test_abi.py).
//!
//! `best_multiexp` / `best_fft` keep the reference's signatures (halo2_proofs/src/arithmetic.rs:132,171). For the two
//! instantiations on the prover's hot path — `C = bn256::G1Affine` and `G = bn256::Fr` — they call libcqb200.so; every
//! other instantiation (G2Affine at static_lookup.rs:146, G1 EC-FFT at arithmetic.rs:285) keeps the generic Rust body,
//! which is a different operation, not a fallback. There is no CPU fallback for the accelerated instantiations: a
//! non-zero return code panics, like the reference's assert!s.
//!
//! Layout contract: `Fr`, `Fq` are `[u64; 4]` Montgomery limbs and `G1Affine` is `{x: Fq, y: Fq}`; the vendored crate must
//! mark them `#[repr(transparent)]` / `#[repr(C)]` (bn256/fr.rs:22-25, derive/curve.rs:163-168), or the shim goes through
//! `SerdeObject::to_raw_bytes` (derive/field.rs:302-308), which yields the same bytes.
use std::any::TypeId;
use std::os::raw::{c_char, c_int, c_void};

use group::Group as _;
use halo2curves::bn256::{Fq, Fr, G1Affine, G1};
use halo2curves::{CurveAffine, Group};

#[allow(non_camel_case_types)]
pub type cqb_bases_t = u64;

extern "C" {
    pub fn cqb_init(device: c_int) -> c_int;
    pub fn cqb_shutdown();
    pub fn cqb_last_error() -> *const c_char;
    pub fn cqb_bases_register(affine_xy: *const u64, n: usize, out: *mut cqb_bases_t) -> c_int;
    pub fn cqb_bases_free(h: cqb_bases_t) -> c_int;
    pub fn cqb_msm_bn254_g1(b: cqb_bases_t, offset: usize, scalars: *const u64, n: usize, out_xy: *mut u64, is_inf: *mut c_int) -> c_int;
    pub fn cqb_msm_bn254_g1_host(affine_xy: *const u64, scalars: *const u64, n: usize, out_xy: *mut u64, is_inf: *mut c_int) -> c_int;
    pub fn cqb_msm_bn254_g1_sparse(b: cqb_bases_t, idx: *const u32, scalars: *const u64, m: usize, out_xy: *mut u64, is_inf: *mut c_int) -> c_int;
    pub fn cqb_bases_register_device(d_affine_xy: *const c_void, n: usize, out: *mut cqb_bases_t) -> c_int;
    pub fn cqb_bases_precompute(h: cqb_bases_t, window_bits: c_int) -> c_int;
    pub fn cqb_msm_bn254_g1_dev(b: cqb_bases_t, offset: usize, d_scalars: *const c_void, n: usize, out_xy: *mut u64, is_inf: *mut c_int) -> c_int;
    pub fn cqb_srs_setup_dev(k: u32, s: *const u64, d_g: *mut c_void, d_g_lagrange: *mut c_void) -> c_int;
    pub fn cqb_table_srs_setup_dev(log_len: u32, s: *const u64, d_g1: *mut c_void, d_g1_lagrange: *mut c_void, d_opening_at_0: *mut c_void) -> c_int;
    pub fn cqb_g_to_lagrange_dev(d_g: *const c_void, k: u32, d_out: *mut c_void) -> c_int;
    pub fn cqb_cq_table_qs_dev(d_table_coeffs: *const c_void, log_n: u32, d_srs_g1: *const c_void, d_qs_out: *mut c_void) -> c_int;
    pub fn cqb_cq_lookup_h_dev(d_values: *mut c_void, d_b_coset: *const c_void, d_f_coset: *const c_void, d_l_active_row: *const c_void,
                               beta: *const u64, y: *const u64, size: u64) -> c_int;
    pub fn cqb_eval_polynomial_dev(d_coeffs: *const c_void, n: usize, point: *const u64, out: *mut u64) -> c_int;
    pub fn cqb_kate_division_dev(d_a: *const c_void, n: usize, b: *const u64, d_q: *mut c_void) -> c_int;
    pub fn cqb_fr_batch_invert_dev(d_a: *mut c_void, n: usize) -> c_int;
    // grand products and the CQ prover's element-wise pieces (device-resident vectors)
    pub fn cqb_fr_prefix_product_dev(d_in: *const c_void, n: usize, init: *const u64, d_out: *mut c_void) -> c_int;
    pub fn cqb_permutation_product_dev(d_columns: *const *const c_void, d_perms: *const *const c_void, ncols: u32, k: u32, beta: *const u64,
                                       gamma: *const u64, omega: *const u64, delta: *const u64, deltaomega_io: *mut u64, last_z: *const u64,
                                       d_z: *mut c_void) -> c_int;
    pub fn cqb_lookup_product_dev(d_compressed_input: *const c_void, d_compressed_table: *const c_void, d_permuted_input: *const c_void,
                                  d_permuted_table: *const c_void, k: u32, beta: *const u64, gamma: *const u64, d_z: *mut c_void) -> c_int;
    pub fn cqb_lookup_h_dev(d_values: *mut c_void, d_table_value: *const c_void, d_product_coset: *const c_void,
                            d_permuted_input_coset: *const c_void, d_permuted_table_coset: *const c_void, d_l0: *const c_void,
                            d_l_last: *const c_void, d_l_active_row: *const c_void, beta: *const u64, gamma: *const u64, y: *const u64,
                            size: u64, rot_scale: i32) -> c_int;
    pub fn cqb_fr_compress_dev(d_cols: *const *const c_void, ncols: u32, d_idx: *const u32, n: usize, theta: *const u64, d_out: *mut c_void) -> c_int;
    pub fn cqb_fr_inv_shifted_dev(d_in: *const c_void, n: usize, usable: usize, shift: *const u64, d_out: *mut c_void) -> c_int;
    pub fn cqb_fr_mul_dev(d_a: *const c_void, d_b: *const c_void, n: usize, d_out: *mut c_void) -> c_int;
    pub fn cqb_msm_bn254_g1_sparse_dev(b: cqb_bases_t, d_idx: *const u32, d_scalars: *const c_void, m: usize, out_xy: *mut u64,
                                       is_inf: *mut c_int) -> c_int;
    // round 2: one process, several GPUs (cqb_init_multi), MSMKZG::eval, the G2 side, batched evaluations
    pub fn cqb_init_multi(n_devices: c_int) -> c_int;
    pub fn cqb_active_devices() -> c_int;
    pub fn cqb_bases_register_sharded(affine_xy: *const u64, n: usize, out: *mut cqb_bases_t) -> c_int;
    pub fn cqb_msm_bn254_g1_multi_dev(b: cqb_bases_t, offset: usize, d_scalars: *const *const c_void, n: usize, out_xy: *mut u64,
                                      is_inf: *mut c_int) -> c_int;
    pub fn cqb_set_host_bases_cache(budget_bytes: i64) -> c_int;
    pub fn cqb_msm_bn254_g1_jacobian(jacobian_xyz: *const u64, scalars: *const u64, n: usize, out_xy: *mut u64, is_inf: *mut c_int) -> c_int;
    pub fn cqb_g1_batch_normalize(jacobian_xyz: *const u64, n: usize, affine_xy_out: *mut u64) -> c_int;
    pub fn cqb_g2_powers(s: *const u64, count: usize, g2_affine_out: *mut u64) -> c_int;
    pub fn cqb_msm_bn254_g2(g2_affine: *const u64, scalars: *const u64, n: usize, out_xy: *mut u64, is_inf: *mut c_int) -> c_int;
    pub fn cqb_eval_polynomials_dev(d_coeffs: *const *const c_void, n: usize, points: *const u64, count: u32, out: *mut u64) -> c_int;
    pub fn cqb_fr_axpy_dev(d_acc: *mut c_void, a: *const u64, d_x: *const c_void, n: usize) -> c_int;
    pub fn cqb_dev_alloc(bytes: usize, d_out: *mut *mut c_void) -> c_int;
    pub fn cqb_dev_free(d: *mut c_void) -> c_int;
    pub fn cqb_memcpy_h2d(d_dst: *mut c_void, h_src: *const c_void, bytes: usize) -> c_int;
    pub fn cqb_memcpy_d2h(h_dst: *mut c_void, d_src: *const c_void, bytes: usize) -> c_int;
    pub fn cqb_ntt_bn254_fr(a: *mut u64, omega: *const u64, log_n: u32) -> c_int;
    pub fn cqb_intt_bn254_fr(a: *mut u64, omega_inv: *const u64, divisor: *const u64, log_n: u32) -> c_int;
    pub fn cqb_coset_ntt_bn254_fr(coeffs: *const u64, n: usize, out: *mut u64, ext_omega: *const u64, ext_log_n: u32,
                                  g_coset: *const u64, g_coset_inv: *const u64) -> c_int;
    pub fn cqb_coset_intt_bn254_fr(a: *mut u64, ext_log_n: u32, ext_omega_inv: *const u64, ext_divisor: *const u64,
                                   g_coset: *const u64, g_coset_inv: *const u64, t_evaluations: *const u64, t_len: u32) -> c_int;
}

fn check(rc: c_int, what: &str) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(cqb_last_error()) }.to_string_lossy().into_owned();
        panic!("{what}: libcqb200 error {rc}: {msg}"); // the reference panics on its assert!s; so does the shim
    }
}

fn g1_from_affine_limbs(xy: [u64; 8], is_inf: c_int) -> G1 {
    if is_inf != 0 {
        return G1::identity();
    }
    // SAFETY: Fq is #[repr(transparent)] over [u64; 4] (layout contract above)
    let x: Fq = unsafe { std::mem::transmute([xy[0], xy[1], xy[2], xy[3]]) };
    let y: Fq = unsafe { std::mem::transmute([xy[4], xy[5], xy[6], xy[7]]) };
    G1 { x, y, z: Fq::one() }
}

/// Replacement for halo2_proofs::arithmetic::best_multiexp (arithmetic.rs:132-159).
pub fn best_multiexp<C: CurveAffine>(coeffs: &[C::Scalar], bases: &[C]) -> C::Curve {
    assert_eq!(coeffs.len(), bases.len()); // arithmetic.rs:133
    if TypeId::of::<C>() == TypeId::of::<G1Affine>() {
        let mut out = [0u64; 8];
        let mut inf: c_int = 0;
        let rc = unsafe {
            cqb_msm_bn254_g1_host(bases.as_ptr() as *const u64, coeffs.as_ptr() as *const u64, coeffs.len(), out.as_mut_ptr(), &mut inf)
        };
        check(rc, "best_multiexp");
        let r = g1_from_affine_limbs(out, inf);
        // SAFETY: C::Curve == G1 was just established through TypeId
        return unsafe { std::mem::transmute_copy::<G1, C::Curve>(&r) };
    }
    generic::best_multiexp(coeffs, bases) // the reference's own body, moved verbatim into `mod generic`
}

/// Replacement for halo2_proofs::arithmetic::best_fft (arithmetic.rs:171-234).
pub fn best_fft<G: Group>(a: &mut [G], omega: G::Scalar, log_n: u32) {
    assert_eq!(a.len(), 1usize << log_n); // arithmetic.rs:184
    if TypeId::of::<G>() == TypeId::of::<Fr>() {
        let rc = unsafe { cqb_ntt_bn254_fr(a.as_mut_ptr() as *mut u64, &omega as *const G::Scalar as *const u64, log_n) };
        check(rc, "best_fft");
        return;
    }
    generic::best_fft(a, omega, log_n)
}

/// ParamsKZG keeps one handle per SRS vector (uploaded once in setup/read); commit / commit_lagrange
/// (poly/kzg/commitment.rs:496-504, 539-543) become:
pub fn commit_with_handle(handle: cqb_bases_t, poly: &[Fr]) -> G1 {
    let mut out = [0u64; 8];
    let mut inf: c_int = 0;
    let rc = unsafe { cqb_msm_bn254_g1(handle, 0, poly.as_ptr() as *const u64, poly.len(), out.as_mut_ptr(), &mut inf) };
    check(rc, "commit");
    g1_from_affine_limbs(out, inf)
}

/// CQ prover: m_cm / a_cm / qa_cm / a0_cm (plonk/static_lookup/prover.rs:167-170, 245-257) as one sparse MSM each.
pub fn commit_sparse(handle: cqb_bases_t, idx: &[u32], scalars: &[Fr]) -> G1 {
    assert_eq!(idx.len(), scalars.len());
    let mut out = [0u64; 8];
    let mut inf: c_int = 0;
    let rc = unsafe { cqb_msm_bn254_g1_sparse(handle, idx.as_ptr(), scalars.as_ptr() as *const u64, idx.len(), out.as_mut_ptr(), &mut inf) };
    check(rc, "commit_sparse");
    g1_from_affine_limbs(out, inf)
}

/// A device-resident vector of `Fr` (cqb_dev_alloc / cqb_memcpy_h2d): the prover keeps its polynomials here between the calls
/// below instead of in `Polynomial<F, B>`'s `Vec`.
#[derive(Copy, Clone)]
pub struct DevFr(pub *mut c_void);

/// One iteration of the column-set loop of `permutation::Argument::commit` (plonk/permutation/prover.rs:82-166): the caller
/// passes the set's columns and `pkey.permutations` (device-resident Lagrange values), gets z in `z`, then writes its random
/// blinding rows into `z[n - blinding_factors..]` (:152-155), reads `last_z = z[n - blinding_factors - 1]` (:157) and commits
/// with `commit_with_handle(g_lagrange_handle, ..)` / `cqb_msm_bn254_g1_dev` (:166). `deltaomega` is updated as at :144.
pub fn permutation_product_set(columns: &[DevFr], permutations: &[DevFr], k: u32, beta: Fr, gamma: Fr, omega: Fr, deltaomega: &mut Fr, last_z: Fr, z: DevFr) {
    assert_eq!(columns.len(), permutations.len());
    let cols: Vec<*const c_void> = columns.iter().map(|c| c.0 as *const c_void).collect();
    let perms: Vec<*const c_void> = permutations.iter().map(|c| c.0 as *const c_void).collect();
    let delta = Fr::DELTA;
    let rc = unsafe {
        cqb_permutation_product_dev(cols.as_ptr(), perms.as_ptr(), cols.len() as u32, k, &beta as *const Fr as *const u64, &gamma as *const Fr as *const u64,
                                    &omega as *const Fr as *const u64, &delta as *const Fr as *const u64, deltaomega as *mut Fr as *mut u64,
                                    &last_z as *const Fr as *const u64, z.0)
    };
    check(rc, "permutation product");
}

/// The device part of `Committed::commit_log_derivatives` (plonk/static_lookup/prover.rs:187-342) up to the sparse
/// commitments: a_i = multiplicity_i / (theta-compressed table value_i + beta) over the support of m, then
/// (a_cm, per-table Q_A parts, a0_cm). `table_values` are the tables' resident value vectors, `qs` their resident cached
/// quotient commitments (StaticTableValues.qs registered with cqb_bases_register_device). The caller combines the Q_A parts
/// with the powers of theta (sum_k theta^(K-1-k) * part_k — the reference's per-index compress_tables, :220-240, by linearity).
pub fn cq_log_derivative_commitments(g1_lagrange: cqb_bases_t, opening_at_0: cqb_bases_t, qs: &[cqb_bases_t], table_values: &[DevFr], support: &[u32],
                                     multiplicities: &[Fr], beta: Fr, theta: Fr, scratch: DevFr /* 3 * |support| Fr + |support| u32 */) -> (G1, Vec<G1>, G1) {
    let m = support.len();
    assert_eq!(m, multiplicities.len());
    unsafe {
        let d_a = scratch.0;
        let d_tv = (scratch.0 as *mut u8).add(m * 32) as *mut c_void;
        let d_mult = (scratch.0 as *mut u8).add(2 * m * 32) as *mut c_void;
        let d_idx = (scratch.0 as *mut u8).add(3 * m * 32) as *mut u32;
        check(cqb_memcpy_h2d(d_idx as *mut c_void, support.as_ptr() as *const c_void, m * 4), "h2d idx");
        check(cqb_memcpy_h2d(d_mult, multiplicities.as_ptr() as *const c_void, m * 32), "h2d multiplicities");
        let tv: Vec<*const c_void> = table_values.iter().map(|t| t.0 as *const c_void).collect();
        check(cqb_fr_compress_dev(tv.as_ptr(), tv.len() as u32, d_idx, m, &theta as *const Fr as *const u64, d_tv), "compress_tables");
        check(cqb_fr_inv_shifted_dev(d_tv, m, m, &beta as *const Fr as *const u64, d_a), "1/(t + beta)");
        check(cqb_fr_mul_dev(d_a, d_mult, m, d_a), "a_i");
        let sparse = |h: cqb_bases_t| {
            let mut out = [0u64; 8];
            let mut inf: c_int = 0;
            check(cqb_msm_bn254_g1_sparse_dev(h, d_idx, d_a, m, out.as_mut_ptr(), &mut inf), "sparse commit");
            g1_from_affine_limbs(out, inf)
        };
        (sparse(g1_lagrange), qs.iter().map(|&h| sparse(h)).collect(), sparse(opening_at_0))
    }
}

mod generic {
    //! The reference's generic bodies of best_multiexp / best_fft (arithmetic.rs:13-159, 171-274) are moved here unchanged
    //! when the shim is applied inside halo2_proofs; they serve the non-accelerated instantiations only.
    use halo2curves::{CurveAffine, Group};
    pub fn best_multiexp<C: CurveAffine>(_coeffs: &[C::Scalar], _bases: &[C]) -> C::Curve {
        unimplemented!("reference body (arithmetic.rs:13-159) lives in halo2_proofs::arithmetic")
    }
    pub fn best_fft<G: Group>(_a: &mut [G], _omega: G::Scalar, _log_n: u32) {
        unimplemented!("reference body (arithmetic.rs:171-274) lives in halo2_proofs::arithmetic")
    }
}

#[allow(dead_code)]
fn _unused(_: *mut c_void) {}

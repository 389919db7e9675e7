// halo2_b200.hpp — C++ host-side mirror of the reference interface for the hot path, over the C ABI of libcqb200.so.
//
// The reference is Rust (no toolchain in this image), so the host layer above the C ABI is written in C++ with the
// reference's names, argument meaning and error behaviour (a reference `assert!` / panic is a thrown std::logic_error):
//   best_multiexp, best_fft, eval_polynomial, kate_division      halo2_proofs/src/arithmetic.rs:132,171,304,351
//   EvaluationDomain::{new, lagrange_to_coeff, coeff_to_extended, divide_by_vanishing_poly, extended_to_coeff}
//                                                                 halo2_proofs/src/poly/domain.rs:39,238,252,319,293
//   ParamsKZG::{setup_from_toxic_waste, commit, commit_lagrange, downsize}
//                                                                 halo2_proofs/src/poly/kzg/commitment.rs:209,539,496,482
//   TableSRS::setup_from_toxic_waste (G1 parts), StaticTableValues::new
//                                                                 poly/kzg/commitment.rs:73-178, plonk/static_lookup.rs:77-126
//   permutation::commit (grand-product sets)                      plonk/permutation/prover.rs:46-200
//   static_lookup::{commit, commit_log_derivatives}               plonk/static_lookup/prover.rs:51-184, 187-342
// Field elements / points use the reference's in-memory layout (4 x u64 Montgomery limbs; affine x||y, identity = zeros).
// Host-side constants (roots of unity, inverses) are computed with the library's own field code (csrc/fp.cuh compiles for the
// host); nothing here touches the test oracle, and there is no CPU fallback for the device operations.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/cqb200.h"
#include "../fp.cuh"

namespace halo2_b200 {

struct Fr {
    uint64_t l[4];
    bool operator==(const Fr& o) const { return std::memcmp(l, o.l, 32) == 0; }
    bool operator!=(const Fr& o) const { return !(*this == o); }
};
struct G1Affine {
    uint64_t x[4], y[4];
    bool is_identity() const { uint64_t o = 0; for (int i = 0; i < 4; i++) o |= x[i] | y[i]; return o == 0; }  // derive/curve.rs:707-709
    bool operator==(const G1Affine& o) const { return std::memcmp(this, &o, 64) == 0; }
};
// C::Curve as returned by best_multiexp; only the affine normal form is canonical
struct G1 {
    G1Affine affine;
    bool identity;
    G1Affine to_affine() const { return affine; }
    bool operator==(const G1& o) const { return affine == o.affine; }
};

/// bn256 G2Affine in the reference's raw layout: x.c0, x.c1, y.c0, y.c1 (Fq Montgomery limbs, 128 bytes); identity = zeros
struct G2Affine {
    uint64_t l[16];
    bool operator==(const G2Affine& o) const { return std::memcmp(l, o.l, 128) == 0; }
};

namespace detail {
inline void check(int rc, const char* what) {
    if (rc != 0) throw std::logic_error(std::string(what) + ": libcqb200 error " + std::to_string(rc) + ": " + cqb_last_error());
}
typedef cqb::Fr F;
inline F to_f(const Fr& a) { F r; for (int i = 0; i < 4; i++) { r.l[2 * i] = (uint32_t)a.l[i]; r.l[2 * i + 1] = (uint32_t)(a.l[i] >> 32); } return r; }
inline Fr from_f(const F& a) { Fr r; for (int i = 0; i < 4; i++) r.l[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32); return r; }
inline F raw(uint64_t a, uint64_t b, uint64_t c, uint64_t d) { Fr t{{a, b, c, d}}; return cqb::fp_to_mont<cqb::FrP>(to_f(t)); }
inline F mul(const F& a, const F& b) { return cqb::fp_mul<cqb::FrP>(a, b); }
inline F inv(const F& a) { return cqb::fp_inv<cqb::FrP>(a); }
inline F pow_u64(F b, uint64_t e) { F r = F::one(); while (e) { if (e & 1) r = mul(r, b); b = mul(b, b); e >>= 1; } return r; }
inline F root_of_unity() { return raw(0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL); }  // fr.rs:77-82
inline F zeta() { return raw(0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL); }           // fr.rs:112-117
}  // namespace detail

inline void init(int device = 0) { detail::check(cqb_init(device), "cqb_init"); }
inline Fr fr_from_u64(uint64_t v) { return detail::from_f(detail::raw(v, 0, 0, 0)); }  // derive/field.rs:114-118 From<u64>
inline Fr fr_one() { return detail::from_f(detail::F::one()); }

/// reference arithmetic.rs:132 — "This function will panic if coeffs and bases have a different length."
inline G1 best_multiexp(const std::vector<Fr>& coeffs, const std::vector<G1Affine>& bases) {
    if (coeffs.size() != bases.size()) throw std::logic_error("assertion failed: `(left == right)` coeffs.len() == bases.len()");  // :133
    G1 r;
    int inf = 0;
    detail::check(cqb_msm_bn254_g1_host((const uint64_t*)bases.data(), (const uint64_t*)coeffs.data(), coeffs.size(), r.affine.x, &inf), "best_multiexp");
    r.identity = inf != 0;
    return r;
}
/// reference arithmetic.rs:171 — in place, n must equal 1 << log_n (:184)
inline void best_fft(std::vector<Fr>& a, const Fr& omega, uint32_t log_n) {
    if (a.size() != ((size_t)1 << log_n)) throw std::logic_error("assertion failed: `(left == right)` n == 1 << log_n");
    detail::check(cqb_ntt_bn254_fr((uint64_t*)a.data(), omega.l, log_n), "best_fft");
}

namespace detail {
struct DevBuf {  // RAII device buffer
    void* p = nullptr;
    explicit DevBuf(size_t bytes) { check(cqb_dev_alloc(bytes, &p), "cqb_dev_alloc"); }
    ~DevBuf() { if (p) cqb_dev_free(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};
}  // namespace detail

/// reference arithmetic.rs:304-329
inline Fr eval_polynomial(const std::vector<Fr>& poly, const Fr& point) {
    detail::DevBuf d(poly.size() * 32 + 32);
    if (!poly.empty()) detail::check(cqb_memcpy_h2d(d.p, poly.data(), poly.size() * 32), "h2d");
    Fr out;
    detail::check(cqb_eval_polynomial_dev(d.p, poly.size(), point.l, out.l), "eval_polynomial");
    return out;
}
/// reference arithmetic.rs:351-387
inline std::vector<Fr> kate_division(const std::vector<Fr>& a, const Fr& b) {
    std::vector<Fr> q(a.empty() ? 0 : a.size() - 1);
    if (q.empty()) return q;
    detail::DevBuf da(a.size() * 32), dq(q.size() * 32);
    detail::check(cqb_memcpy_h2d(da.p, a.data(), a.size() * 32), "h2d");
    detail::check(cqb_kate_division_dev(da.p, a.size(), b.l, dq.p), "kate_division");
    detail::check(cqb_memcpy_d2h(q.data(), dq.p, q.size() * 32), "d2h");
    return q;
}

/// Polynomial<F, ExtendedLagrangeCoeff>; `divided` records a pending divide_by_vanishing_poly (fused into extended_to_coeff)
struct ExtendedLagrange {
    std::vector<Fr> values;
    bool divided = false;
};

/// reference poly/domain.rs:21-34, 39-142
class EvaluationDomain {
  public:
    EvaluationDomain(uint32_t j, uint32_t k) : k_(k), n_((uint64_t)1 << k), quotient_poly_degree_(j - 1) {
        using namespace detail;
        extended_k_ = k;
        while (((uint64_t)1 << extended_k_) < n_ * quotient_poly_degree_) extended_k_++;
        F ew = root_of_unity();
        for (uint32_t i = extended_k_; i < 28; i++) ew = mul(ew, ew);
        F w = ew;
        for (uint32_t i = k; i < extended_k_; i++) w = mul(w, w);
        F gc = zeta(), gci = mul(gc, gc);
        F orig = pow_u64(gc, n_), step = pow_u64(ew, n_), cur = orig;
        do {
            t_evaluations_.push_back(from_f(inv(cqb::fp_sub<cqb::FrP>(cur, F::one()))));  // (t - 1)^-1, batch_invert :118-125
            cur = mul(cur, step);
        } while (!(cur == orig));
        if (t_evaluations_.size() != ((size_t)1 << (extended_k_ - k))) throw std::logic_error("assert_eq!(t_evaluations.len(), 1 << (extended_k - k))");
        omega_ = from_f(w); omega_inv_ = from_f(inv(w));
        extended_omega_ = from_f(ew); extended_omega_inv_ = from_f(inv(ew));
        g_coset_ = from_f(gc); g_coset_inv_ = from_f(gci);
        ifft_divisor_ = from_f(inv(raw((uint64_t)1 << k, 0, 0, 0)));
        extended_ifft_divisor_ = from_f(inv(raw((uint64_t)1 << extended_k_, 0, 0, 0)));
    }
    uint32_t k() const { return k_; }
    uint32_t extended_k() const { return extended_k_; }
    size_t extended_len() const { return (size_t)1 << extended_k_; }
    uint64_t get_quotient_poly_degree() const { return quotient_poly_degree_; }
    const Fr& get_omega() const { return omega_; }
    const Fr& get_omega_inv() const { return omega_inv_; }
    const Fr& get_extended_omega() const { return extended_omega_; }
    const Fr& ifft_divisor() const { return ifft_divisor_; }
    const Fr& get_extended_omega_inv() const { return extended_omega_inv_; }
    const Fr& extended_ifft_divisor() const { return extended_ifft_divisor_; }
    const Fr& g_coset() const { return g_coset_; }
    const Fr& g_coset_inv() const { return g_coset_inv_; }
    const std::vector<Fr>& t_evaluations() const { return t_evaluations_; }

    /// domain.rs:366-374
    static void ifft(std::vector<Fr>& a, const Fr& omega_inv, uint32_t log_n, const Fr& divisor) {
        if (a.size() != ((size_t)1 << log_n)) throw std::logic_error("assertion failed: n == 1 << log_n");
        detail::check(cqb_intt_bn254_fr((uint64_t*)a.data(), omega_inv.l, divisor.l, log_n), "ifft");
    }
    /// domain.rs:238-248
    std::vector<Fr> lagrange_to_coeff(std::vector<Fr> a) const {
        if (a.size() != ((size_t)1 << k_)) throw std::logic_error("assertion failed: a.values.len() == 1 << self.k");
        ifft(a, omega_inv_, k_, ifft_divisor_);
        return a;
    }
    /// domain.rs:252-266
    ExtendedLagrange coeff_to_extended(const std::vector<Fr>& a) const {
        if (a.size() != ((size_t)1 << k_)) throw std::logic_error("assertion failed: a.values.len() == 1 << self.k");
        ExtendedLagrange e;
        e.values.resize(extended_len());
        detail::check(cqb_coset_ntt_bn254_fr((const uint64_t*)a.data(), a.size(), (uint64_t*)e.values.data(), extended_omega_.l, extended_k_,
                                             g_coset_.l, g_coset_inv_.l), "coeff_to_extended");
        return e;
    }
    /// domain.rs:319-338 (recorded; executed fused with extended_to_coeff, as vanishing/prover.rs:84-87 calls them back to back)
    ExtendedLagrange divide_by_vanishing_poly(ExtendedLagrange a) const {
        if (a.values.size() != extended_len()) throw std::logic_error("assertion failed: a.values.len() == self.extended_len()");
        a.divided = true;
        return a;
    }
    /// domain.rs:293-315
    std::vector<Fr> extended_to_coeff(ExtendedLagrange a) const {
        if (a.values.size() != extended_len()) throw std::logic_error("assertion failed: a.values.len() == self.extended_len()");
        detail::check(cqb_coset_intt_bn254_fr((uint64_t*)a.values.data(), extended_k_, extended_omega_inv_.l, extended_ifft_divisor_.l, g_coset_.l,
                                              g_coset_inv_.l, a.divided ? (const uint64_t*)t_evaluations_.data() : nullptr,
                                              a.divided ? (uint32_t)t_evaluations_.size() : 0), "extended_to_coeff");
        a.values.resize((size_t)(n_ * quotient_poly_degree_));
        return a.values;
    }

  private:
    uint32_t k_, extended_k_;
    uint64_t n_, quotient_poly_degree_;
    Fr omega_, omega_inv_, extended_omega_, extended_omega_inv_, g_coset_, g_coset_inv_, ifft_divisor_, extended_ifft_divisor_;
    std::vector<Fr> t_evaluations_;
};

/// reference poly/kzg/commitment.rs:31-39 — the SRS lives in HBM (two cqb_bases_t handles over one device allocation)
class ParamsKZG {
  public:
    /// commitment.rs:209-276 (G1 part): generated on the device
    static ParamsKZG setup_from_toxic_waste(uint32_t k, const Fr& s, bool precompute = true) {
        if (k > 28) throw std::logic_error("assertion failed: k <= E::Scalar::S");
        ParamsKZG p;
        p.k_ = k;
        p.n_ = (uint64_t)1 << k;
        detail::check(cqb_dev_alloc(2 * p.n_ * 64, &p.dev_), "cqb_dev_alloc");
        detail::check(cqb_srs_setup_dev(k, s.l, p.dev_, (char*)p.dev_ + p.n_ * 64), "setup_from_toxic_waste");
        p.register_handles(precompute);
        return p;
    }
    ParamsKZG(ParamsKZG&& o) noexcept { *this = std::move(o); }
    ParamsKZG& operator=(ParamsKZG&& o) noexcept {
        release();
        k_ = o.k_; n_ = o.n_; dev_ = o.dev_; g_ = o.g_; g_lagrange_ = o.g_lagrange_;
        o.dev_ = nullptr; o.g_ = o.g_lagrange_ = 0;
        return *this;
    }
    ~ParamsKZG() { release(); }
    uint32_t k() const { return k_; }
    uint64_t n() const { return n_; }
    /// commitment.rs:539-543
    G1 commit(const std::vector<Fr>& poly) const { return msm(g_, poly); }
    /// commitment.rs:496-504
    G1 commit_lagrange(const std::vector<Fr>& poly) const { return msm(g_lagrange_, poly); }
    /// commitment.rs:482-490
    void downsize(uint32_t k) {
        if (k > k_) throw std::logic_error("assertion failed: k <= self.k");
        uint64_t n = (uint64_t)1 << k;
        void* d = nullptr;
        detail::check(cqb_dev_alloc(2 * n * 64, &d), "cqb_dev_alloc");
        detail::check(cqb_memcpy_d2d(d, dev_, n * 64), "d2d");
        detail::check(cqb_g_to_lagrange_dev(d, k, (char*)d + n * 64), "g_to_lagrange");
        detail::check(cqb_sync(), "sync");
        release();
        k_ = k; n_ = n; dev_ = d;
        register_handles(false);
    }
    std::vector<G1Affine> get_g() const { return download(dev_); }
    std::vector<G1Affine> g_lagrange() const { return download((char*)dev_ + n_ * 64); }
    cqb_bases_t g_handle() const { return g_; }
    cqb_bases_t g_lagrange_handle() const { return g_lagrange_; }

  private:
    ParamsKZG() = default;
    void register_handles(bool precompute) {
        detail::check(cqb_bases_register_device(dev_, n_, &g_), "register g");
        detail::check(cqb_bases_register_device((char*)dev_ + n_ * 64, n_, &g_lagrange_), "register g_lagrange");
        if (precompute && n_ >= ((uint64_t)1 << 16)) {
            detail::check(cqb_bases_precompute(g_, 0), "precompute g");
            detail::check(cqb_bases_precompute(g_lagrange_, 0), "precompute g_lagrange");
        }
    }
    void release() {
        if (g_) cqb_bases_free(g_);
        if (g_lagrange_) cqb_bases_free(g_lagrange_);
        if (dev_) cqb_dev_free(dev_);
        g_ = g_lagrange_ = 0;
        dev_ = nullptr;
    }
    G1 msm(cqb_bases_t h, const std::vector<Fr>& poly) const {
        if (n_ < poly.size()) throw std::logic_error("assertion failed: self.n() >= size as u64");  // :502, :541
        G1 r;
        int inf = 0;
        detail::check(cqb_msm_bn254_g1(h, 0, (const uint64_t*)poly.data(), poly.size(), r.affine.x, &inf), "commit");
        r.identity = inf != 0;
        return r;
    }
    std::vector<G1Affine> download(const void* d) const {
        std::vector<G1Affine> v(n_);
        detail::check(cqb_memcpy_d2h(v.data(), d, n_ * 64), "d2h");
        return v;
    }
    uint32_t k_ = 0;
    uint64_t n_ = 0;
    void* dev_ = nullptr;
    cqb_bases_t g_ = 0, g_lagrange_ = 0;
};

namespace detail {
inline G1 fetch_point(const uint64_t out[8], int inf) {
    G1 r;
    std::memcpy(r.affine.x, out, 64);
    r.identity = inf != 0;
    return r;
}
inline Fr fr_of(const F& a) { return from_f(a); }
inline F delta() { return raw(0x870e56bbe533e9a2ULL, 0x5b5f898e5e963f25ULL, 0x64ec26aad4c86e71ULL, 0x09226b6e22c6f0caULL); }  // fr.rs:87-92
}  // namespace detail

/// reference poly/kzg/commitment.rs:42-47, 73-178 (G1 parts; the G2 powers are keygen / verifier side): device resident
class TableSRS {
  public:
    /// reference signature (:73): also the G2 powers [s^i]G2, i <= max_g2_power (:94-104, 114-141) — keygen / verifier-side data
    static TableSRS setup_from_toxic_waste(size_t max_g1_power, size_t max_g2_power, const Fr& s) {
        TableSRS t = setup_from_toxic_waste(max_g1_power, s);
        t.g2_.resize(max_g2_power + 1);
        detail::check(cqb_g2_powers(s.l, max_g2_power + 1, (uint64_t*)t.g2_.data()), "TableSRS g2 powers");
        return t;
    }
    static TableSRS setup_from_toxic_waste(size_t max_g1_power, const Fr& s) {
        size_t len = max_g1_power + 1;
        if (len & (len - 1)) throw std::logic_error("assertion failed: is_pow_2(g1_len)");  // :77
        uint32_t log_len = 0;
        while (((size_t)1 << log_len) < len) log_len++;
        TableSRS t;
        t.len_ = len;
        detail::check(cqb_dev_alloc(3 * len * 64, &t.dev_), "cqb_dev_alloc");
        detail::check(cqb_table_srs_setup_dev(log_len, s.l, t.dev_, (char*)t.dev_ + len * 64, (char*)t.dev_ + 2 * len * 64), "TableSRS::setup");
        detail::check(cqb_bases_register_device(t.dev_, len, &t.g1_), "register g1");
        detail::check(cqb_bases_register_device((char*)t.dev_ + len * 64, len, &t.g1_lagrange_), "register g1_lagrange");
        detail::check(cqb_bases_register_device((char*)t.dev_ + 2 * len * 64, len, &t.opening_at_0_), "register opening_at_0");
        return t;
    }
    TableSRS(TableSRS&& o) noexcept { *this = std::move(o); }
    TableSRS& operator=(TableSRS&& o) noexcept {
        release();
        len_ = o.len_; dev_ = o.dev_; g1_ = o.g1_; g1_lagrange_ = o.g1_lagrange_; opening_at_0_ = o.opening_at_0_;
        g2_ = std::move(o.g2_);
        o.dev_ = nullptr; o.g1_ = o.g1_lagrange_ = o.opening_at_0_ = 0;
        return *this;
    }
    ~TableSRS() { release(); }
    size_t len() const { return len_; }
    const void* g1_dev() const { return dev_; }
    cqb_bases_t g1() const { return g1_; }
    cqb_bases_t g1_lagrange() const { return g1_lagrange_; }
    cqb_bases_t g_lagrange_opening_at_0() const { return opening_at_0_; }
    const std::vector<G2Affine>& g2() const { return g2_; }
    std::vector<G1Affine> download(int which) const {  // 0: g1, 1: g1_lagrange, 2: g_lagrange_opening_at_0
        std::vector<G1Affine> v(len_);
        detail::check(cqb_memcpy_d2h(v.data(), (char*)dev_ + (size_t)which * len_ * 64, len_ * 64), "d2h");
        return v;
    }

  private:
    TableSRS() = default;
    void release() {
        for (cqb_bases_t h : {g1_, g1_lagrange_, opening_at_0_}) if (h) cqb_bases_free(h);
        if (dev_) cqb_dev_free(dev_);
        dev_ = nullptr; g1_ = g1_lagrange_ = opening_at_0_ = 0;
    }
    size_t len_ = 0;
    void* dev_ = nullptr;
    cqb_bases_t g1_ = 0, g1_lagrange_ = 0, opening_at_0_ = 0;
    std::vector<G2Affine> g2_;
};

/// reference plonk/static_lookup.rs:162-168
struct StaticCommittedTable {
    G2Affine zv, t, x_b0_bound;
    size_t size;
};

/// reference plonk/static_lookup.rs:69-126: a table's values and its cached quotient commitments qs (FK on the device
/// instead of the reference's N kate_divisions + N MSMs, ":107 TODO: THIS SHOULD BE DONE WITH FK METHOD")
class StaticTableValues {
  public:
    StaticTableValues(const std::vector<Fr>& values, const TableSRS& srs) : size_(values.size()), values_(values) {
        if (size_ == 0 || (size_ & (size_ - 1))) throw std::logic_error("assertion failed: is_pow_2(size)");  // :80
        if (srs.len() < size_) throw std::logic_error("srs_g1 shorter than the table");
        for (size_t i = 0; i < size_; i++) {
            std::string key((const char*)values[i].l, 32);
            if (!value_index_mapping_.emplace(key, i).second) throw std::logic_error("table is not all unique values");  // :82-85
        }
        uint32_t k = 0;
        while (((size_t)1 << k) < size_) k++;
        EvaluationDomain dom(2, k);
        detail::check(cqb_dev_alloc(size_ * 32 * 2 + size_ * 64, &dev_), "cqb_dev_alloc");
        void* d_coeffs = dev_;
        d_values_ = (char*)dev_ + size_ * 32;
        d_qs_ = (char*)dev_ + size_ * 64;
        detail::check(cqb_memcpy_h2d(d_coeffs, values.data(), size_ * 32), "h2d");
        detail::check(cqb_memcpy_d2d(d_values_, d_coeffs, size_ * 32), "d2d");
        detail::check(cqb_intt_bn254_fr_dev(d_coeffs, dom.get_omega_inv().l, dom.ifft_divisor().l, k), "table ifft");  // :99-105
        detail::check(cqb_cq_table_qs_dev(d_coeffs, k, srs.g1_dev(), d_qs_), "cq_table_qs");                          // :107-119
        detail::check(cqb_bases_register_device(d_qs_, size_, &qs_), "register qs");
    }
    StaticTableValues(const StaticTableValues&) = delete;
    StaticTableValues& operator=(const StaticTableValues&) = delete;
    ~StaticTableValues() {
        if (qs_) cqb_bases_free(qs_);
        if (dev_) cqb_dev_free(dev_);
    }
    /// reference plonk/static_lookup.rs:127-160: zv = [s^N]G2 - G2, t = best_multiexp::<G2Affine>(ifft(values), srs_g2) with the values
    /// in the order of value_index_mapping.keys() — a BTreeMap, i.e. SORTED by canonical value (derive/field.rs:128-141), as the
    /// reference does — and x_b0_bound = srs_g2[srs_g1_len - 1 - (circuit_domain - 2)]
    StaticCommittedTable commit(size_t srs_g1_len, const std::vector<G2Affine>& srs_g2, size_t circuit_domain) const {
        if (srs_g2.size() <= size_) throw std::logic_error("srs_g2 must hold the powers 0..N");
        std::vector<size_t> order(size_);
        for (size_t i = 0; i < size_; i++) order[i] = i;
        std::vector<detail::F> canon(size_);
        for (size_t i = 0; i < size_; i++) canon[i] = cqb::fp_from_mont<cqb::FrP>(detail::to_f(values_[i]));
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
            for (int w = 7; w >= 0; w--) if (canon[a].l[w] != canon[b].l[w]) return canon[a].l[w] < canon[b].l[w];
            return false;
        });
        std::vector<Fr> coeffs(size_);
        for (size_t i = 0; i < size_; i++) coeffs[i] = values_[order[i]];
        uint32_t k = 0;
        while (((size_t)1 << k) < size_) k++;
        EvaluationDomain dom(2, k);
        EvaluationDomain::ifft(coeffs, dom.get_omega_inv(), k, dom.ifft_divisor());
        StaticCommittedTable r;
        int inf = 0;
        G2Affine two[2] = {srs_g2[size_], srs_g2[0]};
        Fr pm[2] = {fr_one(), detail::from_f(cqb::fp_neg<cqb::FrP>(detail::F::one()))};
        detail::check(cqb_msm_bn254_g2(two[0].l, pm[0].l, 2, r.zv.l, &inf), "zv");
        detail::check(cqb_msm_bn254_g2(srs_g2[0].l, coeffs[0].l, size_, r.t.l, &inf), "table commitment");
        r.x_b0_bound = srs_g2[srs_g1_len - 1 - (circuit_domain - 2)];
        r.size = srs_g1_len;
        return r;
    }
    size_t size() const { return size_; }
    const void* values_dev() const { return d_values_; }
    cqb_bases_t qs() const { return qs_; }
    /// value_index_mapping.get(fi) (:86-89); throws like the reference's `.expect("... not in table")`
    size_t index_of(const Fr& v) const {
        auto it = value_index_mapping_.find(std::string((const char*)v.l, 32));
        if (it == value_index_mapping_.end()) throw std::logic_error("value not in table");
        return it->second;
    }

  private:
    size_t size_;
    std::vector<Fr> values_;
    void* dev_ = nullptr;
    void* d_values_ = nullptr;
    void* d_qs_ = nullptr;
    cqb_bases_t qs_ = 0;
    std::map<std::string, size_t> value_index_mapping_;
};

namespace permutation {
/// One CommittedSet of reference plonk/permutation/prover.rs:23-28 (Lagrange values of z + its commitment)
struct CommittedSet {
    std::vector<Fr> z;
    G1 commitment;
};
/// reference plonk/permutation/prover.rs:46-200 Argument::commit. columns / permutations: the permutation's columns and
/// pkey.permutations in Lagrange form, in self.columns order; blind_rows[set] = the blinding_factors random values the
/// reference draws from its rng for that set (:152-155). Everything is computed on the device.
inline std::vector<CommittedSet> commit(const ParamsKZG& params, const EvaluationDomain& domain, size_t cs_degree, size_t blinding_factors,
                                        const std::vector<std::vector<Fr>>& columns, const std::vector<std::vector<Fr>>& permutations,
                                        const Fr& beta, const Fr& gamma, const std::vector<std::vector<Fr>>& blind_rows) {
    using namespace detail;
    if (cs_degree < 3) throw std::logic_error("assertion failed: pk.vk.cs_degree >= 3");  // :78
    if (columns.size() != permutations.size()) throw std::logic_error("columns / permutations length mismatch");
    const size_t chunk_len = cs_degree - 2, n = params.n(), ncols = columns.size();
    const size_t nsets = (ncols + chunk_len - 1) / chunk_len;
    if (blind_rows.size() != nsets) throw std::logic_error("one row of blinding values per column set is required");
    DevBuf d_cols(std::max<size_t>(1, 2 * ncols) * n * 32), d_z(std::max<size_t>(1, nsets) * n * 32);
    for (size_t j = 0; j < ncols; j++) {
        if (columns[j].size() != n || permutations[j].size() != n) throw std::logic_error("column length != params.n()");
        check(cqb_memcpy_h2d((char*)d_cols.p + j * n * 32, columns[j].data(), n * 32), "h2d");
        check(cqb_memcpy_h2d((char*)d_cols.p + (ncols + j) * n * 32, permutations[j].data(), n * 32), "h2d");
    }
    Fr deltaomega = fr_one(), last_z = fr_one(), delta = from_f(detail::delta());
    std::vector<CommittedSet> sets(nsets);
    for (size_t s = 0; s < nsets; s++) {
        size_t c0 = s * chunk_len, c1 = std::min(ncols, c0 + chunk_len);
        std::vector<const void*> cp, pp;
        for (size_t j = c0; j < c1; j++) { cp.push_back((char*)d_cols.p + j * n * 32); pp.push_back((char*)d_cols.p + (ncols + j) * n * 32); }
        void* z = (char*)d_z.p + s * n * 32;
        check(cqb_permutation_product_dev(cp.data(), pp.data(), (uint32_t)cp.size(), domain.k(), beta.l, gamma.l, domain.get_omega().l, delta.l,
                                          deltaomega.l, last_z.l, z), "permutation product");
        if (blind_rows[s].size() != blinding_factors) throw std::logic_error("blinding row count != blinding_factors");
        if (blinding_factors) check(cqb_memcpy_h2d((char*)z + (n - blinding_factors) * 32, blind_rows[s].data(), blinding_factors * 32), "h2d");  // :152-155
        check(cqb_memcpy_d2h(last_z.l, (char*)z + (n - (blinding_factors + 1)) * 32, 32), "d2h");  // :157
        check(cqb_sync(), "sync");
        uint64_t out[8];
        int inf = 0;
        check(cqb_msm_bn254_g1_dev(params.g_lagrange_handle(), 0, z, n, out, &inf), "commit_lagrange(z)");  // :166
        sets[s].commitment = fetch_point(out, inf);
        sets[s].z.resize(n);
        check(cqb_memcpy_d2h(sets[s].z.data(), z, n * 32), "d2h");
    }
    check(cqb_sync(), "sync");
    return sets;
}
}  // namespace permutation

namespace static_lookup {
/// reference plonk/static_lookup/prover.rs:28-34 Committed
struct Committed {
    std::vector<Fr> f;                 // compressed input expression, Lagrange form
    std::map<size_t, Fr> m_sparse;     // table index -> multiplicity (BTreeMap: key order)
    G1 f_cm, m_cm;
};
/// reference plonk/static_lookup/prover.rs:36-41 CommittedLogDerivative + the commitments written at :301-313
struct CommittedLogDerivative {
    std::vector<Fr> b, b0, f;          // coefficient form
    Fr a_at_zero;
    G1 a_cm, qa_cm, a0_cm, b0_cm, p_cm;
};

/// prover.rs:51-184 Argument::commit for a vector lookup of `inputs` (one evaluated input expression per table, Lagrange form)
inline Committed commit(const ParamsKZG& params, const TableSRS& table_config, const std::vector<const StaticTableValues*>& tables,
                        const std::vector<std::vector<Fr>>& inputs, const Fr& theta, size_t blinding_factors) {
    using namespace detail;
    const size_t n = params.n(), K = tables.size();
    if (K == 0 || inputs.size() != K) throw std::logic_error("one input expression per table is required");
    for (auto* t : tables) if (t->size() != tables[0]->size()) throw std::logic_error("Tables should all be of the same size");  // :82-84
    Committed c;
    DevBuf d_in(K * n * 32), d_f(n * 32);
    std::vector<const void*> ptrs;
    for (size_t j = 0; j < K; j++) {
        if (inputs[j].size() != n) throw std::logic_error("input expression length != params.n()");
        check(cqb_memcpy_h2d((char*)d_in.p + j * n * 32, inputs[j].data(), n * 32), "h2d");
        ptrs.push_back((char*)d_in.p + j * n * 32);
    }
    check(cqb_fr_compress_dev(ptrs.data(), (uint32_t)K, nullptr, n, theta.l, d_f.p), "compress_expressions");  // :108-121
    const size_t usable_rows = n - (blinding_factors + 1);                                                     // :125-126
    for (size_t row = 0; row < usable_rows; row++) {                                                            // :132-160
        bool have = false;
        size_t idx = 0;
        for (size_t j = 0; j < K; j++) {
            size_t index = tables[j]->index_of(inputs[j][row]);
            if (have && idx != index) throw std::logic_error("Vector lookup must be on the same table row");
            idx = index;
            have = true;
        }
        auto it = c.m_sparse.find(idx);
        F one = F::one();
        if (it == c.m_sparse.end()) c.m_sparse[idx] = from_f(one);
        else it->second = from_f(cqb::fp_add<cqb::FrP>(to_f(it->second), one));
    }
    uint64_t out[8];
    int inf = 0;
    check(cqb_msm_bn254_g1_dev(params.g_lagrange_handle(), 0, d_f.p, n, out, &inf), "f_cm");  // :164-165
    c.f_cm = fetch_point(out, inf);
    std::vector<uint32_t> idx;
    std::vector<Fr> mult;
    for (auto& kv : c.m_sparse) { idx.push_back((uint32_t)kv.first); mult.push_back(kv.second); }
    check(cqb_msm_bn254_g1_sparse(table_config.g1_lagrange(), idx.data(), (const uint64_t*)mult.data(), idx.size(), out, &inf), "m_cm");  // :167-170
    c.m_cm = fetch_point(out, inf);
    c.f.resize(n);
    check(cqb_memcpy_d2h(c.f.data(), d_f.p, n * 32), "d2h");
    check(cqb_sync(), "sync");
    return c;
}

/// prover.rs:187-342 Committed::commit_log_derivatives; b0_g1_bound = pk.b0_g1_bound (n - 1 points, :299)
inline CommittedLogDerivative commit_log_derivatives(const Committed& c, const ParamsKZG& params, const EvaluationDomain& domain,
                                                     const TableSRS& table_config, const std::vector<const StaticTableValues*>& tables,
                                                     cqb_bases_t b0_g1_bound, const Fr& beta, const Fr& theta, size_t blinding_factors) {
    using namespace detail;
    const size_t n = params.n(), K = tables.size(), m = c.m_sparse.size(), N = tables[0]->size();
    if (cqb_bases_len(b0_g1_bound) != n - 1) throw std::logic_error("assertion failed: `(left == right)` coeffs.len() == bases.len()");  // arithmetic.rs:133
    std::vector<uint32_t> idx;
    std::vector<Fr> mult;
    for (auto& kv : c.m_sparse) { idx.push_back((uint32_t)kv.first); mult.push_back(kv.second); }
    const size_t mm = std::max<size_t>(m, 1);
    DevBuf d_vec(3 * n * 32), d_sp(3 * mm * 32 + mm * 4);
    void *d_b = d_vec.p, *d_b0 = (char*)d_vec.p + n * 32, *d_f = (char*)d_vec.p + 2 * n * 32;
    void *d_a = d_sp.p, *d_tv = (char*)d_sp.p + mm * 32, *d_mult = (char*)d_sp.p + 2 * mm * 32;
    uint32_t* d_idx = (uint32_t*)((char*)d_sp.p + 3 * mm * 32);
    CommittedLogDerivative r;
    uint64_t out[8];
    int inf = 0;
    auto sparse = [&](cqb_bases_t h) {
        check(cqb_msm_bn254_g1_sparse_dev(h, d_idx, d_a, m, out, &inf), "sparse commit");
        return fetch_point(out, inf);
    };
    if (m) {
        check(cqb_memcpy_h2d(d_idx, idx.data(), m * 4), "h2d");
        check(cqb_memcpy_h2d(d_mult, mult.data(), m * 32), "h2d");
        std::vector<const void*> tv;
        for (auto* t : tables) tv.push_back(t->values_dev());
        check(cqb_fr_compress_dev(tv.data(), (uint32_t)K, d_idx, m, theta.l, d_tv), "compress_tables");   // :224-229
        check(cqb_fr_inv_shifted_dev(d_tv, m, m, beta.l, d_a), "1/(t + beta)");                            // :243
        check(cqb_fr_mul_dev(d_a, d_mult, m, d_a), "a_i");
    }
    r.a_cm = sparse(table_config.g1_lagrange());                                                           // :249
    // :230-240, :250: Q_A over the theta-compressed cached quotients = sum_k theta^(K-1-k) MSM(qs_k, a), by linearity
    std::vector<G1Affine> parts;
    for (auto* t : tables) parts.push_back(sparse(t->qs()).affine);
    if (K == 1) {
        check(cqb_g1_sum_affine((const uint64_t*)parts.data(), 1, out, &inf), "qa");
        r.qa_cm = fetch_point(out, inf);
    } else {
        std::vector<Fr> pw(K);
        F acc = F::one();
        for (size_t j = K; j-- > 0;) { pw[j] = from_f(acc); acc = mul(acc, to_f(theta)); }
        r.qa_cm = best_multiexp(pw, parts);  // K terms; the Python mirror folds theta^j into the scalars instead (cq.py)
    }
    r.a0_cm = sparse(table_config.g_lagrange_opening_at_0());                                              // :252
    const size_t usable_rows = n - (blinding_factors + 1);                                                 // :259-260
    check(cqb_memcpy_h2d(d_f, c.f.data(), n * 32), "h2d");
    check(cqb_fr_inv_shifted_dev(d_f, n, usable_rows, beta.l, d_b), "bs");                                 // :261-269
    check(cqb_intt_bn254_fr_dev(d_b, domain.get_omega_inv().l, domain.ifft_divisor().l, domain.k()), "ifft(bs)");  // :271-276
    check(cqb_memcpy_d2d(d_b0, (char*)d_b + 32, (n - 1) * 32), "b0");                                      // :279
    Fr zero{{0, 0, 0, 0}};
    check(cqb_memcpy_h2d((char*)d_b0 + (n - 1) * 32, zero.l, 32), "b0 push zero");                         // :303
    check(cqb_msm_bn254_g1_dev(b0_g1_bound, 0, d_b0, n - 1, out, &inf), "p_cm");                           // :299
    r.p_cm = fetch_point(out, inf);
    check(cqb_msm_bn254_g1_dev(params.g_handle(), 0, d_b0, n, out, &inf), "b0_cm");                        // :310
    r.b0_cm = fetch_point(out, inf);
    check(cqb_intt_bn254_fr_dev(d_f, domain.get_omega_inv().l, domain.ifft_divisor().l, domain.k()), "ifft(f)");  // :327-332
    r.b.resize(n); r.b0.resize(n); r.f.resize(n);
    check(cqb_memcpy_d2h(r.b.data(), d_b, n * 32), "d2h");
    check(cqb_memcpy_d2h(r.b0.data(), d_b0, n * 32), "d2h");
    check(cqb_memcpy_d2h(r.f.data(), d_f, n * 32), "d2h");
    check(cqb_sync(), "sync");
    // :315-325 A(0) = (n B(0) - (blinding_factors + 1) / beta) / N, B(0) = b's constant coefficient
    F b_at_zero = to_f(r.b[0]), beta_inv = inv(to_f(beta));
    F t = cqb::fp_sub<cqb::FrP>(mul(b_at_zero, raw(n, 0, 0, 0)), mul(raw(blinding_factors + 1, 0, 0, 0), beta_inv));
    r.a_at_zero = from_f(mul(t, inv(raw(N, 0, 0, 0))));
    return r;
}
}  // namespace static_lookup

/// reference poly/kzg/msm.rs:12-80: scalars and PROJECTIVE bases collected term by term; eval() = batch_normalize + best_multiexp
struct G1Jacobian { uint64_t x[4], y[4], z[4]; };
class MSMKZG {
  public:
    void append_term(const Fr& scalar, const G1Jacobian& point) { scalars_.push_back(scalar); bases_.push_back(point); }  // :41-44
    void add_msm(const MSMKZG& other) {                                                                                    // :46-49
        scalars_.insert(scalars_.end(), other.scalars_.begin(), other.scalars_.end());
        bases_.insert(bases_.end(), other.bases_.begin(), other.bases_.end());
    }
    void scale(const Fr& factor) {                                                                                         // :51-58
        for (auto& sc : scalars_) sc = detail::from_f(detail::mul(detail::to_f(sc), detail::to_f(factor)));
    }
    G1 eval() const {                                                                                                      // :65-70
        uint64_t out[8];
        int inf = 0;
        detail::check(cqb_msm_bn254_g1_jacobian(bases_.empty() ? nullptr : bases_[0].x, scalars_.empty() ? nullptr : scalars_[0].l,
                                                scalars_.size(), out, &inf), "MSMKZG::eval");
        return detail::fetch_point(out, inf);
    }
    bool check() const { return eval().identity; }                                                                      // :60-62
    const std::vector<Fr>& scalars() const { return scalars_; }
    const std::vector<G1Jacobian>& bases() const { return bases_; }

  private:
    std::vector<Fr> scalars_;
    std::vector<G1Jacobian> bases_;
};

}  // namespace halo2_b200

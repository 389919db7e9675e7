// halo2_b200.hpp — C++ host-side mirror of the reference interface for the hot path, over the C ABI of libcqb200.so.
//
// The reference is Rust (no toolchain in this image), so the host layer above the C ABI is written in C++ with the
// reference's names, argument meaning and error behaviour (a reference `assert!` / panic is a thrown std::logic_error):
//   best_multiexp, best_fft, eval_polynomial, kate_division      halo2_proofs/src/arithmetic.rs:132,171,304,351
//   EvaluationDomain::{new, lagrange_to_coeff, coeff_to_extended, divide_by_vanishing_poly, extended_to_coeff}
//                                                                 halo2_proofs/src/poly/domain.rs:39,238,252,319,293
//   ParamsKZG::{setup_from_toxic_waste, commit, commit_lagrange, downsize}
//                                                                 halo2_proofs/src/poly/kzg/commitment.rs:209,539,496,482
// Field elements / points use the reference's in-memory layout (4 x u64 Montgomery limbs; affine x||y, identity = zeros).
// Host-side constants (roots of unity, inverses) are computed with the library's own field code (csrc/fp.cuh compiles for the
// host); nothing here touches the test oracle, and there is no CPU fallback for the device operations.
#pragma once
#include <cstdint>
#include <cstring>
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

}  // namespace halo2_b200

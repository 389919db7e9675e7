// halo2_b200_prover.hpp — C++ mirror of halo2_proofs::plonk::create_proof (reference halo2_proofs/src/plonk/prover.rs:37-797) with
// the KZG / GWC backend (poly/kzg/multiopen/gwc/prover.rs:42-86), for circuits made of advice columns, one permutation argument and
// static (CQ) lookups — the call path of the sha crate's CQ circuits. Same steps, same order of transcript writes as the reference
// (and as the Python mirror sha2-on-cq-halo2_b200/prover.py, whose proof bytes the tests compare these with); every polynomial stays
// in HBM from the witness upload to the last opening witness (one pooled device allocation per proving key). The transcript is the
// caller's (the reference's Blake2bWrite, transcript.rs:199-240): this header only calls it in the reference's order.
#pragma once
#include <utility>

#include "halo2_b200.hpp"

namespace halo2_b200 {
namespace plonk {

/// transcript.rs:35-62 TranscriptWrite (+ Transcript::common_scalar / squeeze_challenge_scalar)
struct Transcript {
    virtual void common_scalar(const Fr& s) = 0;
    virtual void write_point(const G1Affine& p) = 0;
    virtual void write_scalar(const Fr& s) = 0;
    virtual Fr squeeze_challenge_scalar() = 0;
    virtual ~Transcript() {}
};

/// one lookup_static of the constraint system (plonk/static_lookup.rs:170-191) with what keygen prepared for it
struct StaticLookup {
    std::vector<size_t> input_columns;             // advice columns whose theta-compression is looked up
    const TableSRS* table_srs;                     // poly/kzg/commitment.rs:42-47
    std::vector<const StaticTableValues*> tables;  // plonk/static_lookup.rs:69-126, in table_ids order
    cqb_bases_t b0_g1_bound;                       // pk.b0_g1_bound: n - 1 points
};

/// m_sparse of one lookup in key order (static_lookup/prover.rs:123-160 builds it on the CPU)
struct SparseM {
    std::vector<uint32_t> idx;
    std::vector<Fr> mult;
};

/// the values the reference draws from its RngCore
struct ProofRng {
    std::vector<std::vector<Fr>> permutation_blinds;  // per column set, blinding_factors values (permutation/prover.rs:152-155)
    std::vector<Fr> random_poly;                      // vanishing/prover.rs:46-55
};

namespace field {
using detail::F;
inline F f(const Fr& a) { return detail::to_f(a); }
inline Fr r(const F& a) { return detail::from_f(a); }
inline F add(const F& a, const F& b) { return cqb::fp_add<cqb::FrP>(a, b); }
inline F sub(const F& a, const F& b) { return cqb::fp_sub<cqb::FrP>(a, b); }
inline F mul(const F& a, const F& b) { return cqb::fp_mul<cqb::FrP>(a, b); }
inline F inv(const F& a) { return cqb::fp_inv<cqb::FrP>(a); }
inline F u64(uint64_t v) { return detail::raw(v, 0, 0, 0); }
inline F pow(F b, uint64_t e) { return detail::pow_u64(b, e); }
}  // namespace field

/// The device-resident part of plonk::ProvingKey that create_proof reads (plonk/keygen.rs:300-400)
class ProvingKey {
  public:
    ProvingKey(const ParamsKZG& params, uint32_t k, size_t cs_degree, size_t blinding_factors, std::vector<size_t> permutation_columns,
               const std::vector<std::vector<Fr>>& sigma_lagrange, std::vector<std::pair<size_t, int>> advice_queries,
               std::vector<StaticLookup> static_lookups, const Fr& vk_transcript_repr)
        : params_(params), k_(k), n_((size_t)1 << k), cs_degree_(cs_degree), bf_(blinding_factors), domain_((uint32_t)cs_degree, k),
          perm_cols_(std::move(permutation_columns)), advice_queries_(std::move(advice_queries)), lookups_(std::move(static_lookups)),
          vk_repr_(vk_transcript_repr) {
        const size_t n = n_, en = domain_.extended_len();
        auto coeff_and_coset = [&](const std::vector<Fr>& lagrange, void** d_coeff, void** d_coset) {
            *d_coeff = alloc(n * 32);
            detail::check(cqb_memcpy_h2d(*d_coeff, lagrange.data(), n * 32), "h2d");
            detail::check(cqb_intt_bn254_fr_dev(*d_coeff, domain_.get_omega_inv().l, domain_.ifft_divisor().l, k_), "ifft");
            *d_coset = alloc(en * 32);
            detail::check(cqb_coset_ntt_bn254_fr_dev(*d_coeff, n, *d_coset, domain_.get_extended_omega().l, domain_.extended_k(),
                                                     domain_.g_coset().l, domain_.g_coset_inv().l), "coeff_to_extended");
        };
        for (const auto& sg : sigma_lagrange) {  // permutation::ProvingKey { permutations, polys, cosets }
            void* dl = alloc(n * 32);
            detail::check(cqb_memcpy_h2d(dl, sg.data(), n * 32), "h2d");
            sigma_lagrange_.push_back(dl);
            void *dc, *de;
            coeff_and_coset(sg, &dc, &de);
            sigma_polys_.push_back(dc);
            sigma_cosets_.push_back(de);
        }
        // l0, l_blind, l_last (keygen.rs:344-363), l_active_row = 1 - (l_last + l_blind) on the extended domain (:367-373)
        const Fr one = fr_one();
        std::vector<Fr> lag(n, Fr{{0, 0, 0, 0}});
        void *dc, *d_lblind;
        lag[0] = one;
        coeff_and_coset(lag, &dc, &l0_);
        lag[0] = Fr{{0, 0, 0, 0}};
        for (size_t i = n - bf_; i < n; i++) lag[i] = one;
        coeff_and_coset(lag, &dc, &d_lblind);
        for (size_t i = n - bf_; i < n; i++) lag[i] = Fr{{0, 0, 0, 0}};
        lag[n - bf_ - 1] = one;
        coeff_and_coset(lag, &dc, &l_last_);
        std::vector<Fr> ones(en, one);
        l_active_row_ = alloc(en * 32);
        detail::check(cqb_memcpy_h2d(l_active_row_, ones.data(), en * 32), "h2d");
        const Fr minus_one = field::r(cqb::fp_neg<cqb::FrP>(detail::F::one()));
        detail::check(cqb_fr_axpy_dev(d_lblind, one.l, l_last_, en), "l_last + l_blind");
        detail::check(cqb_fr_axpy_dev(d_lblind, minus_one.l, l_active_row_, en), "1 - (l_last + l_blind)");
        detail::check(cqb_memcpy_d2d(l_active_row_, d_lblind, en * 32), "d2d");
        detail::check(cqb_sync(), "sync");
    }
    ProvingKey(const ProvingKey&) = delete;
    ProvingKey& operator=(const ProvingKey&) = delete;
    ~ProvingKey() {
        for (void* p : owned_) cqb_dev_free(p);
        if (pool_) cqb_dev_free(pool_);
    }
    /// poly/domain.rs:414-424
    Fr rotate_omega(const Fr& x, int rot) const {
        using namespace field;
        F w = rot >= 0 ? pow(f(domain_.get_omega()), (uint64_t)rot) : pow(f(domain_.get_omega_inv()), (uint64_t)(-rot));
        return r(mul(f(x), w));
    }

  private:
    friend void create_proof(ProvingKey&, const std::vector<std::vector<Fr>>&, const std::vector<SparseM>&, const ProofRng&, Transcript&);
    void* alloc(size_t bytes) {
        void* p = nullptr;
        detail::check(cqb_dev_alloc(bytes ? bytes : 64, &p), "cqb_dev_alloc");
        owned_.push_back(p);
        return p;
    }
    // the working memory of one create_proof, allocated once and reused by every proof
    void ensure_pool(size_t n_advice) {
        const size_t n = n_, en = domain_.extended_len(), chunk = cs_degree_ - 2;
        const size_t nsets = perm_cols_.empty() ? 0 : (perm_cols_.size() + chunk - 1) / chunk, L = lookups_.size();
        const size_t elems = n * (2 * n_advice + 2 * nsets + 8 * L + 6) + en * (nsets + n_advice + 2 * L + 1);
        const size_t need = elems * 32 + ((size_t)1 << 20);
        if (need > pool_bytes_) {
            if (pool_) cqb_dev_free(pool_);
            pool_ = nullptr;
            detail::check(cqb_dev_alloc(need, &pool_), "proof pool");
            pool_bytes_ = need;
        }
    }
    const ParamsKZG& params_;
    uint32_t k_;
    size_t n_, cs_degree_, bf_;
    EvaluationDomain domain_;
    std::vector<size_t> perm_cols_;
    std::vector<std::pair<size_t, int>> advice_queries_;
    std::vector<StaticLookup> lookups_;
    Fr vk_repr_;
    std::vector<void*> owned_, sigma_lagrange_, sigma_polys_, sigma_cosets_;
    void *l0_ = nullptr, *l_last_ = nullptr, *l_active_row_ = nullptr;
    void* pool_ = nullptr;
    size_t pool_bytes_ = 0;
};

/// plonk/prover.rs:37-797 for ONE circuit instance. advice: Lagrange values per column (blinding rows already drawn by the caller).
inline void create_proof(ProvingKey& pk, const std::vector<std::vector<Fr>>& advice, const std::vector<SparseM>& lookups_m_sparse,
                         const ProofRng& rng, Transcript& transcript) {
    using namespace field;
    using detail::check;
    const EvaluationDomain& dom = pk.domain_;
    const ParamsKZG& params = pk.params_;
    const uint32_t k = pk.k_;
    const size_t n = pk.n_, en = dom.extended_len(), bf = pk.bf_, A = advice.size();
    pk.ensure_pool(A);
    size_t used = 0;
    auto alloc = [&](size_t bytes) -> void* {  // a pointer bump inside the pool
        size_t start = (used + 255) & ~(size_t)255;
        if (start + bytes > pk.pool_bytes_) throw std::logic_error("proof pool exhausted");
        used = start + bytes;
        return (char*)pk.pool_ + start;
    };
    auto at = [](void* p, size_t elems) -> void* { return (char*)p + elems * 32; };
    auto to_coeff = [&](const void* d_lagrange) {
        void* d = alloc(n * 32);
        check(cqb_memcpy_d2d(d, d_lagrange, n * 32), "d2d");
        check(cqb_intt_bn254_fr_dev(d, dom.get_omega_inv().l, dom.ifft_divisor().l, k), "lagrange_to_coeff");
        return d;
    };
    auto to_extended = [&](const void* d_coeff) {
        void* d = alloc(en * 32);
        check(cqb_coset_ntt_bn254_fr_dev(d_coeff, n, d, dom.get_extended_omega().l, dom.extended_k(), dom.g_coset().l, dom.g_coset_inv().l),
              "coeff_to_extended");
        return d;
    };
    auto commit_batch = [&](cqb_bases_t h, const void* d, size_t count, size_t batch) {
        std::vector<G1Affine> out(batch);
        std::vector<int> inf(batch);
        for (size_t done = 0; done < batch; done += 64) {
            const int b = (int)std::min<size_t>(64, batch - done);
            check(cqb_msm_bn254_g1_batch_dev(h, 0, (const char*)d + done * count * 32, count, b, out[done].x, &inf[done]), "commit batch");
        }
        return out;
    };
    auto commit = [&](cqb_bases_t h, const void* d, size_t count) {
        G1Affine out;
        int inf = 0;
        check(cqb_msm_bn254_g1_dev(h, 0, d, count, out.x, &inf), "commit");
        return out;
    };

    transcript.common_scalar(pk.vk_repr_);                                                  // prover.rs:85
    void* d_adv0 = alloc(A * n * 32);
    std::vector<void*> d_adv(A);
    for (size_t i = 0; i < A; i++) {
        if (advice[i].size() != n) throw std::logic_error("advice column of the wrong length");
        d_adv[i] = at(d_adv0, i * n);
        check(cqb_memcpy_h2d(d_adv[i], advice[i].data(), n * 32), "h2d");
    }
    for (const auto& pt : commit_batch(params.g_lagrange_handle(), d_adv0, n, A)) transcript.write_point(pt);  // :356-374
    const Fr theta = transcript.squeeze_challenge_scalar();                                 // :472
    // static lookups, first phase: f and m (static_lookup/prover.rs:51-184)
    std::vector<void*> d_f;
    for (size_t l = 0; l < pk.lookups_.size(); l++) {
        const StaticLookup& lk = pk.lookups_[l];
        const SparseM& ms = lookups_m_sparse[l];
        std::vector<const void*> cols;
        for (size_t c : lk.input_columns) cols.push_back(d_adv[c]);
        void* d = alloc(n * 32);
        check(cqb_fr_compress_dev(cols.data(), (uint32_t)cols.size(), nullptr, n, theta.l, d), "compress inputs");  // :108-121
        d_f.push_back(d);
        transcript.write_point(commit(params.g_lagrange_handle(), d, n));                   // f_cm :165, :174
        G1Affine m_cm;
        int inf = 0;
        check(cqb_msm_bn254_g1_sparse(lk.table_srs->g1_lagrange(), ms.idx.data(), ms.mult.empty() ? nullptr : ms.mult[0].l, ms.idx.size(),
                                      m_cm.x, &inf), "m_cm");                               // :167-175
        transcript.write_point(m_cm);
    }
    const Fr beta = transcript.squeeze_challenge_scalar();                                  // :529
    const Fr gamma = transcript.squeeze_challenge_scalar();                                 // :532
    // permutation argument (permutation/prover.rs:46-200)
    const size_t chunk_len = pk.cs_degree_ - 2, ncols = pk.perm_cols_.size();
    const size_t nsets = ncols ? (ncols + chunk_len - 1) / chunk_len : 0;
    void* d_z0 = alloc(std::max<size_t>(nsets, 1) * n * 32);
    std::vector<void*> d_z(nsets), z_poly, z_coset;
    if (nsets) {
        Fr deltaomega = fr_one(), last_z = fr_one();
        const Fr delta = r(detail::delta());
        for (size_t s_ = 0; s_ < nsets; s_++) {                                             // :82-163
            d_z[s_] = at(d_z0, s_ * n);
            std::vector<const void*> cols, perms;
            for (size_t j = s_ * chunk_len; j < std::min(ncols, (s_ + 1) * chunk_len); j++) {
                cols.push_back(d_adv[pk.perm_cols_[j]]);
                perms.push_back(pk.sigma_lagrange_[j]);
            }
            check(cqb_permutation_product_dev(cols.data(), perms.data(), (uint32_t)cols.size(), k, beta.l, gamma.l, dom.get_omega().l, delta.l,
                                              deltaomega.l, last_z.l, d_z[s_]), "permutation product");
            if (bf) check(cqb_memcpy_h2d(at(d_z[s_], n - bf), rng.permutation_blinds[s_].data(), bf * 32), "blinding rows");  // :152-155
            check(cqb_memcpy_d2h(last_z.l, at(d_z[s_], n - bf - 1), 32), "last_z");         // :157
            check(cqb_sync(), "sync");
        }
        for (const auto& pt : commit_batch(params.g_lagrange_handle(), d_z0, n, nsets)) transcript.write_point(pt);  // :166-186
        for (size_t s_ = 0; s_ < nsets; s_++) z_poly.push_back(to_coeff(d_z[s_]));          // :168
        for (size_t s_ = 0; s_ < nsets; s_++) z_coset.push_back(to_extended(z_poly[s_]));   // :171
    }
    // static lookups, second phase (static_lookup/prover.rs:187-342), every vector resident
    struct Cld { void *d_b, *d_b0, *d_f; Fr a_at_zero; size_t N; };
    std::vector<Cld> clds;
    for (size_t l = 0; l < pk.lookups_.size(); l++) {
        const StaticLookup& lk = pk.lookups_[l];
        const SparseM& ms = lookups_m_sparse[l];
        const size_t m = ms.idx.size(), K = lk.tables.size(), N = lk.tables[0]->size(), usable = n - (bf + 1);
        Cld c;
        c.N = N;
        c.d_b = alloc(n * 32);
        c.d_b0 = alloc(n * 32);
        c.d_f = alloc(n * 32);
        void* d_a = alloc(std::max<size_t>(m, 1) * 32);
        void* d_tv = alloc(std::max<size_t>(m, 1) * 32);
        void* d_mult = alloc(std::max<size_t>(m, 1) * 32);
        void* d_idx = alloc(std::max<size_t>(m, 1) * 4);
        G1Affine pt;
        int inf = 0;
        auto sparse = [&](cqb_bases_t h, const void* d_sc) {
            check(cqb_msm_bn254_g1_sparse_dev(h, (const uint32_t*)d_idx, d_sc, m, pt.x, &inf), "sparse commitment");
            return pt;
        };
        if (m) {
            check(cqb_memcpy_h2d(d_idx, ms.idx.data(), m * 4), "h2d");
            check(cqb_memcpy_h2d(d_mult, ms.mult.data(), m * 32), "h2d");
            std::vector<const void*> tv;
            for (auto* t : lk.tables) tv.push_back(t->values_dev());
            check(cqb_fr_compress_dev(tv.data(), (uint32_t)K, (const uint32_t*)d_idx, m, theta.l, d_tv), "compress_tables");  // :224-229
            check(cqb_fr_inv_shifted_dev(d_tv, m, m, beta.l, d_a), "1/(t + beta)");                                             // :243
            check(cqb_fr_mul_dev(d_a, d_mult, m, d_a), "a_i");
        }
        const G1Affine a_cm = sparse(lk.table_srs->g1_lagrange(), d_a);                     // :249
        // :230-240, :250: Q_A over the theta-compressed cached quotients = sum_j MSM(qs_j, theta^(K-1-j) a), by linearity
        std::vector<G1Affine> parts;
        F pw = F::one();
        std::vector<F> pws(K);
        for (size_t j = K; j-- > 0;) { pws[j] = pw; pw = mul(pw, f(theta)); }
        for (size_t j = 0; j < K; j++) {
            if (pws[j] == F::one()) { parts.push_back(sparse(lk.tables[j]->qs(), d_a)); continue; }
            check(cqb_memcpy_d2d(d_tv, d_a, m * 32), "d2d");
            check(cqb_fr_scale_dev(d_tv, m, r(pws[j]).l), "theta power");
            parts.push_back(sparse(lk.tables[j]->qs(), d_tv));
        }
        G1Affine qa_cm;
        check(cqb_g1_sum_affine(parts[0].x, K, qa_cm.x, &inf), "Q_A");
        const G1Affine a0_cm = sparse(lk.table_srs->g_lagrange_opening_at_0(), d_a);        // :252
        check(cqb_fr_inv_shifted_dev(d_f[l], n, usable, beta.l, c.d_b), "bs");              // :261-269
        check(cqb_intt_bn254_fr_dev(c.d_b, dom.get_omega_inv().l, dom.ifft_divisor().l, k), "ifft(bs)");  // :271-276
        check(cqb_memcpy_d2d(c.d_b0, at(c.d_b, 1), (n - 1) * 32), "b0");                    // :279
        const Fr zero{{0, 0, 0, 0}};
        check(cqb_memcpy_h2d(at(c.d_b0, n - 1), zero.l, 32), "b0 push zero");               // :303
        const G1Affine p_cm = commit(lk.b0_g1_bound, c.d_b0, n - 1);                        // :299
        const G1Affine b0_cm = commit(params.g_handle(), c.d_b0, n);                        // :310
        for (const G1Affine& q : {a_cm, qa_cm, a0_cm, b0_cm, p_cm}) transcript.write_point(q);  // :301-313
        Fr b_at_zero;
        check(cqb_memcpy_d2h(b_at_zero.l, c.d_b, 32), "B(0)");
        check(cqb_sync(), "sync");
        // :315-325 A(0) = (n B(0) - (blinding_factors + 1) / beta) / N
        c.a_at_zero = r(mul(sub(mul(f(b_at_zero), u64(n)), mul(u64(bf + 1), inv(f(beta)))), inv(u64(N))));
        check(cqb_memcpy_d2d(c.d_f, d_f[l], n * 32), "f");                                  // :327-334
        check(cqb_intt_bn254_fr_dev(c.d_f, dom.get_omega_inv().l, dom.ifft_divisor().l, k), "ifft(f)");
        clds.push_back(c);
    }
    // vanishing argument: random polynomial (vanishing/prover.rs:37-65)
    void* d_rnd = alloc(n * 32);
    check(cqb_memcpy_h2d(d_rnd, rng.random_poly.data(), n * 32), "h2d");
    transcript.write_point(commit(params.g_handle(), d_rnd, n));
    const Fr y = transcript.squeeze_challenge_scalar();                                     // prover.rs:584
    // advice polys and h(X) (prover.rs:587-624, evaluation.rs:285-551)
    std::vector<void*> adv_poly(A), adv_coset(A);
    for (size_t i = 0; i < A; i++) adv_poly[i] = to_coeff(d_adv[i]);
    for (size_t i = 0; i < A; i++) adv_coset[i] = to_extended(adv_poly[i]);
    void* d_h = alloc(en * 32);
    {
        std::vector<Fr> zeros(en, Fr{{0, 0, 0, 0}});
        check(cqb_memcpy_h2d(d_h, zeros.data(), en * 32), "h2d");
    }
    const int32_t rot_scale = (int32_t)1 << (dom.extended_k() - k);
    if (nsets) {
        std::vector<const void*> sets(z_coset.begin(), z_coset.end()), cols, perms(pk.sigma_cosets_.begin(), pk.sigma_cosets_.end());
        for (size_t c : pk.perm_cols_) cols.push_back(adv_coset[c]);
        check(cqb_permutation_h_dev(d_h, en, rot_scale, -(int32_t)(bf + 1), (uint32_t)chunk_len, sets.data(), (uint32_t)nsets, cols.data(),
                                    perms.data(), (uint32_t)ncols, pk.l0_, pk.l_last_, pk.l_active_row_, beta.l, gamma.l, y.l,
                                    dom.get_extended_omega().l), "permutation terms");
    }
    for (const Cld& c : clds) {                                                             // evaluation.rs:533-548
        void* b_coset = to_extended(c.d_b);
        void* f_coset = to_extended(c.d_f);
        check(cqb_cq_lookup_h_dev(d_h, b_coset, f_coset, pk.l_active_row_, beta.l, y.l, en), "CQ term");
    }
    // vanishing construct (vanishing/prover.rs:69-120)
    check(cqb_coset_intt_bn254_fr_dev(d_h, dom.extended_k(), dom.get_extended_omega_inv().l, dom.extended_ifft_divisor().l, dom.g_coset().l,
                                      dom.g_coset_inv().l, (const uint64_t*)dom.t_evaluations().data(), (uint32_t)dom.t_evaluations().size()),
          "divide + extended_to_coeff");
    const size_t npieces = (size_t)dom.get_quotient_poly_degree();
    for (const auto& pt : commit_batch(params.g_handle(), d_h, n, npieces)) transcript.write_point(pt);
    const Fr x = transcript.squeeze_challenge_scalar();                                     // prover.rs:627
    const F xn = pow(f(x), n);
    // h(X) = sum_i h_i(X) xn^i (vanishing/prover.rs:131-135)
    void* d_hx = alloc(n * 32);
    check(cqb_memcpy_d2d(d_hx, at(d_h, (npieces - 1) * n), n * 32), "d2d");
    for (size_t i = npieces - 1; i-- > 0;) check(cqb_fr_axpy_dev(d_hx, r(xn).l, at(d_h, i * n), n), "h(X)");
    // every evaluation, queued at once, read back together; written in the reference's order
    const Fr x_next = pk.rotate_omega(x, 1), x_last = pk.rotate_omega(x, -(int)(bf + 1));
    std::vector<const void*> ev_p;
    std::vector<Fr> ev_x;
    auto q = [&](const void* p, const Fr& pt) { ev_p.push_back(p); ev_x.push_back(pt); };
    for (const auto& aq : pk.advice_queries_) q(adv_poly[aq.first], pk.rotate_omega(x, aq.second));  // prover.rs:652-670
    q(d_rnd, x);                                                                            // vanishing/prover.rs:137-138
    for (void* p : pk.sigma_polys_) q(p, x);                                                // permutation/prover.rs:229-241
    for (size_t s_ = 0; s_ < nsets; s_++) {                                                 // :244-288
        q(z_poly[s_], x);
        q(z_poly[s_], x_next);
        if (s_ + 1 < nsets) q(z_poly[s_], x_last);
    }
    for (const Cld& c : clds) { q(c.d_b0, x); q(c.d_f, x); }                                // static_lookup/prover.rs:346-375
    q(d_hx, x);
    std::vector<Fr> ev(ev_p.size());
    check(cqb_eval_polynomials_dev(ev_p.data(), n, ev_x[0].l, (uint32_t)ev_p.size(), ev[0].l), "evaluations");
    size_t pos = 0;
    struct Query { Fr point; const void* poly; Fr eval; };
    std::vector<Query> queries;
    for (const auto& aq : pk.advice_queries_) {
        transcript.write_scalar(ev[pos]);
        queries.push_back({pk.rotate_omega(x, aq.second), adv_poly[aq.first], ev[pos]});
        pos++;
    }
    const Fr random_eval = ev[pos++];
    transcript.write_scalar(random_eval);
    std::vector<Fr> sigma_evals;
    for (size_t j = 0; j < pk.sigma_polys_.size(); j++) { sigma_evals.push_back(ev[pos]); transcript.write_scalar(ev[pos++]); }
    std::vector<Fr> z_cur(nsets), z_nxt(nsets), z_lst(nsets);
    for (size_t s_ = 0; s_ < nsets; s_++) {
        z_cur[s_] = ev[pos++];
        z_nxt[s_] = ev[pos++];
        transcript.write_scalar(z_cur[s_]);
        transcript.write_scalar(z_nxt[s_]);
        if (s_ + 1 < nsets) { z_lst[s_] = ev[pos++]; transcript.write_scalar(z_lst[s_]); }
    }
    std::vector<std::pair<Fr, Fr>> lk_evals;
    for (const Cld& c : clds) {
        const Fr b0_eval = ev[pos++], f_eval = ev[pos++];
        transcript.write_scalar(b0_eval);
        transcript.write_scalar(f_eval);
        transcript.write_scalar(c.a_at_zero);
        lk_evals.push_back({b0_eval, f_eval});
    }
    const Fr h_eval = ev[pos++];
    // the queries in the order prover.rs:718-774 chains them
    for (size_t s_ = 0; s_ < nsets; s_++) {                                                 // permutation/prover.rs:304-340
        queries.push_back({x, z_poly[s_], z_cur[s_]});
        queries.push_back({x_next, z_poly[s_], z_nxt[s_]});
    }
    for (size_t s_ = nsets; s_-- > 0;) if (s_ + 1 < nsets) queries.push_back({x_last, z_poly[s_], z_lst[s_]});
    for (size_t l = 0; l < clds.size(); l++) {                                              // static_lookup/prover.rs:378-400
        queries.push_back({x, clds[l].d_b0, lk_evals[l].first});
        queries.push_back({x, clds[l].d_f, lk_evals[l].second});
    }
    for (size_t j = 0; j < pk.sigma_polys_.size(); j++) queries.push_back({x, pk.sigma_polys_[j], sigma_evals[j]});  // pk.permutation.open
    queries.push_back({x, d_hx, h_eval});                                                   // vanishing/prover.rs:160-173
    queries.push_back({x, d_rnd, random_eval});
    // GWC multi-open (poly/kzg/multiopen/gwc/prover.rs:42-86)
    const Fr v = transcript.squeeze_challenge_scalar();
    std::vector<std::pair<Fr, std::vector<Query>>> point_sets;                              // construct_intermediate_sets (gwc.rs:36-60)
    for (const Query& qu : queries) {
        bool found = false;
        for (auto& ps : point_sets)
            if (std::memcmp(ps.first.l, qu.point.l, 32) == 0) { ps.second.push_back(qu); found = true; break; }
        if (!found) point_sets.push_back({qu.point, {qu}});
    }
    void* d_batch = alloc(n * 32);
    void* d_wit0 = alloc(point_sets.size() * n * 32);
    {
        std::vector<Fr> zeros(point_sets.size() * n, Fr{{0, 0, 0, 0}});
        check(cqb_memcpy_h2d(d_wit0, zeros.data(), zeros.size() * 32), "h2d");
    }
    for (size_t j = 0; j < point_sets.size(); j++) {
        const auto& qs = point_sets[j].second;
        check(cqb_memcpy_d2d(d_batch, qs.back().poly, n * 32), "d2d");                      // sum_i v^i p_i by Horner from the last query
        F eval_batch = f(qs.back().eval);
        for (size_t i = qs.size() - 1; i-- > 0;) {
            check(cqb_fr_axpy_dev(d_batch, v.l, qs[i].poly, n), "poly batch");
            eval_batch = add(mul(eval_batch, f(v)), f(qs[i].eval));
        }
        Fr c0;
        check(cqb_memcpy_d2h(c0.l, d_batch, 32), "c0");
        check(cqb_sync(), "sync");
        c0 = r(sub(f(c0), eval_batch));                                                     // poly_batch - eval_batch: the constant coefficient
        check(cqb_memcpy_h2d(d_batch, c0.l, 32), "c0");
        check(cqb_kate_division_dev(d_batch, n, point_sets[j].first.l, at(d_wit0, j * n)), "kate_division");  // arithmetic.rs:351-387
    }
    for (const auto& pt : commit_batch(params.g_handle(), d_wit0, n, point_sets.size())) transcript.write_point(pt);  // gwc/prover.rs:79-84
    check(cqb_sync(), "sync");
}

}  // namespace plonk
}  // namespace halo2_b200

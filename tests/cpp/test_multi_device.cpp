// One process, several GPUs, nothing but the C ABI (include/cqb200.h): what a single-process Rust prover binds.
// cqb_init_multi(n) -> cqb_bases_register_sharded -> cqb_msm_bn254_g1 from pageable and pinned host scalars, offsets and short
// MSMs, cqb_msm_bn254_g1_multi_dev with resident scalars — each result equal to the CPU oracle's best_multiexp
// (halo2_proofs/src/arithmetic.rs:132-159) and to the single-device result. Built and run by tests/test_gpu_multi_device.py.
// usage: test_multi_device <n_devices> <log_n>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/cqb200.h"

extern "C" {
void oracle_synth_scalars(uint64_t seed, size_t start, size_t n, uint64_t* out);
void oracle_synth_bases(uint64_t seed, size_t n, size_t threads, uint64_t* out);
void oracle_best_multiexp(const uint64_t* coeffs, const uint64_t* bases, size_t len, size_t num_threads, uint64_t* out_jac, uint64_t* out_aff);
int oracle_hw_threads(void);
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        int rc_ = (call);                                                                                 \
        if (rc_ != 0) { fprintf(stderr, "FAILED %s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #call, rc_, cqb_last_error()); exit(1); } \
    } while (0)
#define ASSERT(c) do { if (!(c)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

static void shard_range(size_t n, int i, int g, size_t* start, size_t* cnt) {
    size_t base = n / g, rem = n % g;
    *start = i * base + ((size_t)i < rem ? i : rem);
    *cnt = base + ((size_t)i < rem ? 1 : 0);
}

int main(int argc, char** argv) {
    int want = argc > 1 ? atoi(argv[1]) : 2;
    int log_n = argc > 2 ? atoi(argv[2]) : 18;
    int have = cqb_device_count();
    if (have < 1) { fprintf(stderr, "no CUDA device\n"); return 2; }
    int g = want < have ? want : have;
    const size_t n = ((size_t)1 << log_n) + 12345;  // not a multiple of the device count
    const int threads = oracle_hw_threads();
    std::vector<uint64_t> bases(n * 8), scalars(n * 4);
    oracle_synth_bases(0xC0FFEE, n, threads, bases.data());
    oracle_synth_scalars(0x5EED0001, 0, n, scalars.data());
    memset(&scalars[4 * 5], 0, 32);        // a zero scalar
    memset(&bases[8 * 7], 0, 64);          // an identity base
    uint64_t jac[12], exp[8], exp_off[8], exp_short[8];
    oracle_best_multiexp(scalars.data(), bases.data(), n, threads, jac, exp);
    const size_t off = n / 3 + 1, cnt_off = n - off - 77;  // a slice that starts and ends inside shards
    oracle_best_multiexp(scalars.data(), bases.data() + off * 8, cnt_off, threads, jac, exp_off);
    const size_t n_short = 1000;           // lives in the first shard only
    oracle_best_multiexp(scalars.data(), bases.data(), n_short, threads, jac, exp_short);

    CK(cqb_init_multi(g));
    ASSERT(cqb_active_devices() == g);
    cqb_bases_t h = 0;
    CK(cqb_bases_register_sharded(bases.data(), n, &h));
    ASSERT(cqb_bases_len(h) == n);
    uint64_t out[8];
    int inf = -1;
    for (int pass = 0; pass < 2; pass++) {  // windowed layout, then the per-shard precomputed tables
        if (pass == 1) {
            CK(cqb_bases_precompute(h, 0));
            ASSERT(cqb_bases_precomputed_window_bits(h) > 0);
        }
        CK(cqb_msm_bn254_g1(h, 0, scalars.data(), n, out, &inf));  // pageable host memory
        ASSERT(memcmp(out, exp, 64) == 0 && inf == 0);
        CK(cqb_msm_bn254_g1(h, off, scalars.data(), cnt_off, out, &inf));
        ASSERT(memcmp(out, exp_off, 64) == 0);
        CK(cqb_msm_bn254_g1(h, 0, scalars.data(), n_short, out, &inf));
        ASSERT(memcmp(out, exp_short, 64) == 0);
        void* pin = nullptr;
        CK(cqb_host_alloc_pinned(n * 32, &pin));
        memcpy(pin, scalars.data(), n * 32);
        CK(cqb_msm_bn254_g1(h, 0, (const uint64_t*)pin, n, out, &inf));  // pinned host memory
        ASSERT(memcmp(out, exp, 64) == 0);
        CK(cqb_host_free_pinned(pin));
        // resident scalars: one buffer per device holding that shard's range
        std::vector<void*> d_sc(g, nullptr);
        for (int i = 0; i < g; i++) {
            size_t s0, c0;
            shard_range(n, i, g, &s0, &c0);
            CK(cqb_dev_alloc_on(i, c0 * 32, &d_sc[i]));
            CK(cqb_memcpy_h2d_on(i, d_sc[i], scalars.data() + s0 * 4, c0 * 32));
        }
        CK(cqb_msm_bn254_g1_multi_dev(h, 0, d_sc.data(), n, out, &inf));
        ASSERT(memcmp(out, exp, 64) == 0);
        for (int i = 0; i < g; i++) CK(cqb_dev_free_on(i, d_sc[i]));
    }
    // download walks the shards
    std::vector<uint64_t> back(n * 8);
    CK(cqb_bases_download(h, 0, n, back.data()));
    ASSERT(memcmp(back.data(), bases.data(), n * 64) == 0);
    // an all-zero scalar vector gives the identity
    std::vector<uint64_t> zeros(n * 4, 0);
    CK(cqb_msm_bn254_g1(h, 0, zeros.data(), n, out, &inf));
    ASSERT(inf == 1);
    CK(cqb_bases_free(h));
    // single-device entry points still work on the primary device after multi init
    cqb_bases_t h1 = 0;
    CK(cqb_bases_register(bases.data(), n_short, &h1));
    CK(cqb_msm_bn254_g1(h1, 0, scalars.data(), n_short, out, &inf));
    ASSERT(memcmp(out, exp_short, 64) == 0);
    CK(cqb_bases_free(h1));
    cqb_shutdown();
    printf("ALL OK devices=%d n=%zu launches=%llu\n", g, n, cqb_launch_count());
    return 0;
}

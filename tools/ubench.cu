// ubench.cu — integer-pipe calibration for the B200 roofline ("SM count and IMAD issue rate must be
// calibrated by a dependent-free IMAD microbenchmark on the box"). Standalone: nvcc -o ubench ubench.cu ; ./ubench
// Prints one JSON object per line: {"test":..., "ops_per_clk_per_sm":..., "gops":..., "sm_mhz_eff":...}
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../sha2-on-cq-halo2_b200/csrc/fp.cuh"
using namespace cqb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent chains per thread

struct Res { unsigned long long cyc; unsigned int sink; };

#define LOOP_BODY(STMT)                                                \
    unsigned long long t0 = clock64();                                 \
    for (int it = 0; it < ITERS; it++) {                               \
        _Pragma("unroll") for (int k = 0; k < CH; k++) { STMT; }       \
    }                                                                  \
    unsigned long long t1 = clock64();

__global__ void k_imad_lo(Res* out, uint32_t s) {
    uint32_t a[CH], x = s | 1u, y = s * 7u + 3u;
    for (int k = 0; k < CH; k++) a[k] = threadIdx.x + k;
    LOOP_BODY(asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(x), "r"(y)))
    uint32_t z = 0; for (int k = 0; k < CH; k++) z ^= a[k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 0x12345) out[blockIdx.x].sink = z;
}
__global__ void k_imad_hi(Res* out, uint32_t s) {
    uint32_t a[CH], x = s | 0x80000001u, y = s * 7u + 3u;
    for (int k = 0; k < CH; k++) a[k] = 0x9e3779b9u * (threadIdx.x + k + 1);
    LOOP_BODY(asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(x), "r"(y)))
    uint32_t z = 0; for (int k = 0; k < CH; k++) z ^= a[k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 0x12345) out[blockIdx.x].sink = z;
}
__global__ void k_imad_wide(Res* out, uint32_t s) {
    unsigned long long a[CH]; uint32_t x = s | 0x80000001u, y = s * 7u + 3u;
    for (int k = 0; k < CH; k++) a[k] = 0x9e3779b9ull * (threadIdx.x + k + 1);
    LOOP_BODY(asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"((uint32_t)a[(k + 1) % CH]), "r"(y)); (void)x)
    unsigned long long z = 0; for (int k = 0; k < CH; k++) z ^= a[k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 0x12345) out[blockIdx.x].sink = (uint32_t)z;
}
// wide with multiplicand dependent on the running value (defeats any operand-reuse cache)
__global__ void k_imad_wide_dep(Res* out, uint32_t s) {
    unsigned long long a[CH]; uint32_t y = s * 7u + 3u;
    for (int k = 0; k < CH; k++) a[k] = 0x9e3779b9ull * (threadIdx.x + k + 1);
    LOOP_BODY(asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(a[k]) : "r"(y)))
    unsigned long long z = 0; for (int k = 0; k < CH; k++) z ^= a[k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 0x12345) out[blockIdx.x].sink = (uint32_t)z;
}
// carry-chained wide MACs: 4 x (mad.lo.cc, madc.hi.cc) = one 8-limb row; counts 4 wide products per row
__global__ void k_row_chain(Res* out, uint32_t s) {
    uint32_t acc[2][9], x[8]; uint32_t y = s * 7u + 3u;
    for (int k = 0; k < 9; k++) { acc[0][k] = threadIdx.x + k; acc[1][k] = threadIdx.x * 3 + k; }
    for (int k = 0; k < 8; k++) x[k] = 0x9e3779b9u * (threadIdx.x + k + s);
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        row_mad(acc[0], x, y);
        row_mad(acc[1], x + 1, y);
        y += acc[0][0];
    }
    unsigned long long t1 = clock64();
    uint32_t z = 0; for (int k = 0; k < 9; k++) z ^= acc[0][k] ^ acc[1][k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 0x12345) out[blockIdx.x].sink = z;
}
// the same chain with a run-time trip count: calibration runs of >= 20 ms, long enough for the SM clock to settle where a
// real MSM runs (the sub-millisecond tests above read 1.2-1.9 GHz effective); cycles / event time = the clock sample
__global__ void k_row_chain_long(Res* out, uint32_t s, int iters) {
    uint32_t acc[2][9], x[8]; uint32_t y = s * 7u + 3u;
    for (int k = 0; k < 9; k++) { acc[0][k] = threadIdx.x + k; acc[1][k] = threadIdx.x * 3 + k; }
    for (int k = 0; k < 8; k++) x[k] = 0x9e3779b9u * (threadIdx.x + k + s);
    unsigned long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
        row_mad(acc[0], x, y);
        row_mad(acc[1], x + 1, y);
        y += acc[0][0];
    }
    unsigned long long t1 = clock64();
    uint32_t z = 0; for (int k = 0; k < 9; k++) z ^= acc[0][k] ^ acc[1][k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 0x12345) out[blockIdx.x].sink = z;
}
// four independent row accumulators per thread, multiplier word fixed: no dependency between rows at all
__global__ void k_row_chain4_long(Res* out, uint32_t s, int iters) {
    uint32_t acc[4][9], x[8]; uint32_t y = s * 7u + 3u;
    for (int r = 0; r < 4; r++) for (int k = 0; k < 9; k++) acc[r][k] = threadIdx.x * (r + 1) + k;
    for (int k = 0; k < 8; k++) x[k] = 0x9e3779b9u * (threadIdx.x + k + s);
    unsigned long long t0 = clock64();
#pragma unroll 2
    for (int it = 0; it < iters; it++) {
        row_mad(acc[0], x, y);
        row_mad(acc[1], x + 1, y);
        row_mad(acc[2], x, y);
        row_mad(acc[3], x + 1, y);
    }
    unsigned long long t1 = clock64();
    uint32_t z = 0; for (int r = 0; r < 4; r++) for (int k = 0; k < 9; k++) z ^= acc[r][k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 0x12345) out[blockIdx.x].sink = z;
}
__global__ void k_iadd3(Res* out, uint32_t s) {
    uint32_t a[CH], x = s | 1u;
    for (int k = 0; k < CH; k++) a[k] = threadIdx.x + k;
    LOOP_BODY(asm volatile("add.u32 %0, %0, %1;" : "+r"(a[k]) : "r"(x)))
    uint32_t z = 0; for (int k = 0; k < CH; k++) z ^= a[k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 0x12345) out[blockIdx.x].sink = z;
}
// 1 wide MAC + 1 add per slot: do they co-issue?
__global__ void k_wide_plus_add(Res* out, uint32_t s) {
    unsigned long long a[CH]; uint32_t b[CH]; uint32_t x = s | 0x80000001u, y = s * 7u + 3u;
    for (int k = 0; k < CH; k++) { a[k] = 0x9e3779b9ull * (threadIdx.x + k + 1); b[k] = k; }
    LOOP_BODY(asm volatile("mad.wide.u32 %0, %2, %3, %0; add.u32 %1, %1, %4;" : "+l"(a[k]), "+r"(b[k]) : "r"((uint32_t)a[(k + 1) % CH]), "r"(y), "r"(x)))
    unsigned long long z = 0; for (int k = 0; k < CH; k++) z ^= a[k] ^ b[k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 0x12345) out[blockIdx.x].sink = (uint32_t)z;
}
// FP64 pipe: dependent-free DFMA chains
__global__ void k_dfma(Res* out, uint32_t s) {
    double a[CH], x = 1.0000001 + s * 1e-12, y = 0.5 + s * 1e-13;
    for (int k = 0; k < CH; k++) a[k] = threadIdx.x + k;
    LOOP_BODY(asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[k]) : "d"(x), "d"(y)))
    double z = 0; for (int k = 0; k < CH; k++) z += a[k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 12345.678) out[blockIdx.x].sink = 1;
}
// DFMA and IMAD.WIDE-row interleaved: do the FP64 pipe and the integer multiplier run concurrently?
__global__ void k_dfma_plus_row(Res* out, uint32_t s) {
    double a[CH], x = 1.0000001 + s * 1e-12, y = 0.5 + s * 1e-13;
    for (int k = 0; k < CH; k++) a[k] = threadIdx.x + k;
    uint32_t acc[2][9], xx[8]; uint32_t yy = s * 7u + 3u;
    for (int k = 0; k < 9; k++) { acc[0][k] = threadIdx.x + k; acc[1][k] = threadIdx.x * 3 + k; }
    for (int k = 0; k < 8; k++) xx[k] = 0x9e3779b9u * (threadIdx.x + k + s);
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        row_mad(acc[0], xx, yy);
#pragma unroll
        for (int k = 0; k < CH; k++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[k]) : "d"(x), "d"(y));
        row_mad(acc[1], xx + 1, yy);
        yy += acc[0][0];
    }
    unsigned long long t1 = clock64();
    double z = 0; for (int k = 0; k < CH; k++) z += a[k];
    uint32_t zi = 0; for (int k = 0; k < 9; k++) zi ^= acc[0][k] ^ acc[1][k];
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    if (z == 12345.678 || zi == 0x12345) out[blockIdx.x].sink = 1;
}
// full Montgomery multiplications, NCHAIN independent dependent-chains per thread
template <class P, int NCHAIN>
__global__ void k_fpmul(Res* out, const Fp<P>* in, Fp<P>* o, int iters) {
    Fp<P> x[NCHAIN], y = in[threadIdx.x];
    for (int k = 0; k < NCHAIN; k++) x[k] = in[threadIdx.x + 32 * (k + 1)];
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < NCHAIN; k++) x[k] = fp_mul<P>(x[k], y);
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    Fp<P> acc = x[0];
    for (int k = 1; k < NCHAIN; k++) acc = fp_add<P>(acc, x[k]);
    o[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// safegcd inversions (fp_inv_safegcd), NCHAIN independent dependent-chains per thread: latency and throughput of the
// branch-free inversion the batched-affine accumulation runs once per thread per step
template <int NCHAIN>
__global__ void k_fqinv(Res* out, const Fq* in, Fq* o, int iters) {
    Fq x[NCHAIN], y = in[threadIdx.x];
    for (int k = 0; k < NCHAIN; k++) x[k] = in[threadIdx.x + 32 * (k + 1)];
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < NCHAIN; k++) x[k] = fp_add<FqP>(fp_inv_safegcd<FqP>(x[k]), y);
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x].cyc = t1 - t0;
    Fq acc = x[0];
    for (int k = 1; k < NCHAIN; k++) acc = fp_add<FqP>(acc, x[k]);
    o[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
static void run(const char* name, F launch, double ops_per_thread, int blocks, int threads, int nsm) {
    Res* d; CK(cudaMalloc(&d, sizeof(Res) * blocks)); CK(cudaMemset(d, 0, sizeof(Res) * blocks));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(d);  // warm
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0)); launch(d); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    std::vector<Res> h(blocks); CK(cudaMemcpy(h.data(), d, sizeof(Res) * blocks, cudaMemcpyDeviceToHost));
    double cyc = 0; for (auto& r : h) cyc += (double)r.cyc; cyc /= blocks;
    double total_ops = ops_per_thread * (double)blocks * threads;
    double waves = (double)blocks / nsm;  // blocks resident at once per SM = blocks/nsm (we size so all are co-resident)
    double ops_clk_sm = ops_per_thread * threads * waves / cyc;
    printf("{\"test\":\"%s\",\"blocks\":%d,\"threads\":%d,\"ms\":%.4f,\"gops\":%.1f,\"avg_block_cycles\":%.0f,\"ops_per_clk_per_sm\":%.2f,\"sm_mhz_eff\":%.0f}\n",
           name, blocks, threads, best, total_ops / best / 1e6, cyc, ops_clk_sm, cyc / best / 1e3);
    fflush(stdout);
    CK(cudaFree(d));
}

int main(int argc, char** argv) {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int nsm = prop.multiProcessorCount;
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d,\"cc\":\"%d.%d\"}\n", prop.name, nsm, prop.clockRate, prop.major, prop.minor);
    uint32_t s = (uint32_t)(argc > 1 ? atoi(argv[1]) : 12345);
    const double per = (double)ITERS * CH;
    for (int bps : {1, 2, 4}) {
        int blocks = nsm * bps, th = 256;
        run("imad_lo", [&](Res* d) { k_imad_lo<<<blocks, th>>>(d, s); }, per, blocks, th, nsm);
        run("imad_hi", [&](Res* d) { k_imad_hi<<<blocks, th>>>(d, s); }, per, blocks, th, nsm);
        run("imad_wide", [&](Res* d) { k_imad_wide<<<blocks, th>>>(d, s); }, per, blocks, th, nsm);
        run("imad_wide_dep", [&](Res* d) { k_imad_wide_dep<<<blocks, th>>>(d, s); }, per, blocks, th, nsm);
        run("row_chain_wideX", [&](Res* d) { k_row_chain<<<blocks, th>>>(d, s); }, (double)ITERS * 8, blocks, th, nsm);
        run("iadd3", [&](Res* d) { k_iadd3<<<blocks, th>>>(d, s); }, per, blocks, th, nsm);
        run("wide_plus_add(pairs)", [&](Res* d) { k_wide_plus_add<<<blocks, th>>>(d, s); }, per, blocks, th, nsm);
        run("dfma", [&](Res* d) { k_dfma<<<blocks, th>>>(d, s); }, per, blocks, th, nsm);
        run("dfma(8)+row_wideX(8)", [&](Res* d) { k_dfma_plus_row<<<blocks, th>>>(d, s); }, (double)ITERS * 8, blocks, th, nsm);
    }
    // long calibration runs (>= 20 ms each): the roofline denominator bench.py uses comes from these lines
    for (int bps : {2, 4}) {
        int blocks = nsm * bps, th = 256;
        for (int iters : {100000, 400000}) {
            char nm[64];
            snprintf(nm, 64, "row_chain_wideX_long_%dk", iters / 1000);
            run(nm, [&](Res* d) { k_row_chain_long<<<blocks, th>>>(d, s, iters); }, (double)iters * 8, blocks, th, nsm);
        }
    }
    for (int bps : {1, 2, 4}) {
        int blocks = nsm * bps, th = 256, iters = 200000;
        char nm[64];
        snprintf(nm, 64, "row_chain4_wideX_long_%dk", iters / 1000);
        run(nm, [&](Res* d) { k_row_chain4_long<<<blocks, th>>>(d, s, iters); }, (double)iters * 16, blocks, th, nsm);
    }
    // modmul throughput
    std::vector<uint32_t> hin(8 * 32 * 16);
    for (size_t i = 0; i < hin.size(); i++) hin[i] = (uint32_t)(0x9e3779b9u * (i + 1)) >> ((i % 8 == 7) ? 3 : 0);
    Fq* din; CK(cudaMalloc(&din, hin.size() * 4)); CK(cudaMemcpy(din, hin.data(), hin.size() * 4, cudaMemcpyHostToDevice));
    Fq* dout; CK(cudaMalloc(&dout, sizeof(Fq) * nsm * 16 * 1024));
    for (int th : {32, 128}) for (int bps : {1, 2, 4}) {  // inversion latency (1 warp per SM) and throughput
        int blocks = nsm * bps;
        char nm[64];
        snprintf(nm, 64, "fq_inv_safegcd_chain1_%dthr", th);
        run(nm, [&](Res* d) { k_fqinv<1><<<blocks, th>>>(d, din, dout, 50); }, 50.0, blocks, th, nsm);
        snprintf(nm, 64, "fq_inv_safegcd_chain2_%dthr", th);
        run(nm, [&](Res* d) { k_fqinv<2><<<blocks, th>>>(d, din, dout, 50); }, 100.0, blocks, th, nsm);
    }
    {   // long modmul runs: 256 threads x 4 CTAs/SM, two chains
        int blocks = nsm * 4, th = 256;
        for (int itl : {20000, 80000}) {
            char nm[64];
            snprintf(nm, 64, "fq_mul_chain2_long_%dk", itl / 1000);
            run(nm, [&](Res* d) { k_fpmul<FqP, 2><<<blocks, th>>>(d, din, dout, itl); }, (double)itl * 2, blocks, th, nsm);
        }
    }
    const int it = 2000;
    for (int th : {128, 256, 512}) for (int bps : {1, 2, 4}) {
        if (th * bps > 1024) continue;
        int blocks = nsm * bps;
        char nm[64];
        snprintf(nm, 64, "fq_mul_chain1"); run(nm, [&](Res* d) { k_fpmul<FqP, 1><<<blocks, th>>>(d, din, dout, it); }, (double)it * 1, blocks, th, nsm);
        snprintf(nm, 64, "fq_mul_chain2"); run(nm, [&](Res* d) { k_fpmul<FqP, 2><<<blocks, th>>>(d, din, dout, it); }, (double)it * 2, blocks, th, nsm);
        snprintf(nm, 64, "fq_mul_chain4"); run(nm, [&](Res* d) { k_fpmul<FqP, 4><<<blocks, th>>>(d, din, dout, it); }, (double)it * 4, blocks, th, nsm);
    }
    return 0;
}

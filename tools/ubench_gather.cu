// tools/ubench_gather.cu — how many DRAM bytes does a random 32 B / 64 B read of a 64 B record cost on B200, per load flavour?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_gather tools/ubench_gather.cu ; run under ncu with
// --metrics gpu__time_duration.sum,dram__bytes_read.sum
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int MODE>
__device__ __forceinline__ uint4 ld(const uint4* p) {
    uint4 r;
    if (MODE == 0) r = __ldg(p);
    else if (MODE == 1) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (MODE == 4) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// each thread reads `bytes` (32 or 64) of one random 64 B record
template <int MODE, int BYTES>
__global__ void gather(const uint4* tab, size_t nrec, size_t nreads, uint4* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nreads) return;
    uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    const uint4* p = tab + (h % nrec) * 4;
    uint4 a = ld<MODE>(p), b = ld<MODE>(p + 1);
    uint32_t s = a.x ^ b.y;
    if (BYTES == 64) { uint4 c = ld<MODE>(p + 2), d = ld<MODE>(p + 3); s ^= c.z ^ d.w; }
    if (s == 0x12345678u) out[i & 1023] = a;
}
// 4 lanes read one 64 B record together (one coalesced 64 B request per record instead of four 16 B requests by its owner)
template <int LANES>
__global__ void gather_coop(const uint4* tab, size_t nrec, size_t nreads, uint4* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    const unsigned long long mine = (unsigned long long)(tab + (h % nrec) * 4);
    const int lane = threadIdx.x & 31;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < LANES; k++) {
        const int owner = lane / LANES + (32 / LANES) * k;
        const uint4* p = (const uint4*)__shfl_sync(0xffffffffu, mine, owner);
        uint4 a = __ldg(p + (lane % LANES));
        s ^= a.x ^ a.y;
    }
    if (s == 0x12345678u) out[i & 1023] = make_uint4(s, 0, 0, 0);
}
int main(int argc, char** argv) {
    const size_t nrec = (size_t)1 << (argc > 1 ? atoi(argv[1]) : 27), nreads = (size_t)1 << 26;
    printf("{\"table_MB\": %zu}\n", nrec * 64 >> 20);
    uint4 *tab, *out;
    cudaMalloc(&tab, nrec * 64);
    cudaMalloc(&out, 1024 * 16);
    cudaMemset(tab, 1, nrec * 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned grid = (unsigned)(nreads / 256);
#define RUN(M, B)                                                                   \
    for (int rep = 0; rep < 2; rep++) {                                             \
        cudaEventRecord(e0);                                                        \
        gather<M, B><<<grid, 256>>>(tab, nrec, nreads, out);                        \
        cudaEventRecord(e1); cudaEventSynchronize(e1);                              \
        float ms; cudaEventElapsedTime(&ms, e0, e1);                                \
        if (rep) printf("{\"mode\": %d, \"bytes\": %d, \"ms\": %.3f, \"useful_GBs\": %.1f}\n", M, B, ms, nreads * (double)B / ms / 1e6); \
    }
    RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(5, 32)
    RUN(0, 64) RUN(1, 64) RUN(2, 64) RUN(3, 64) RUN(4, 64) RUN(5, 64)
#define RUNC(LN)                                                                    \
    for (int rep = 0; rep < 2; rep++) {                                             \
        cudaEventRecord(e0);                                                        \
        gather_coop<LN><<<grid, 256>>>(tab, nrec, nreads, out);                     \
        cudaEventRecord(e1); cudaEventSynchronize(e1);                              \
        float ms; cudaEventElapsedTime(&ms, e0, e1);                                \
        if (rep) printf("{\"coop_lanes\": %d, \"bytes\": %d, \"ms\": %.3f, \"useful_GBs\": %.1f}\n", LN, LN * 16, ms, nreads * (double)(LN * 16) / ms / 1e6); \
    }
    RUNC(2) RUNC(4) RUNC(8)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// gen.cu — synthetic benchmark/test inputs generated on the device; definitions mirror
// oracle_synth_scalars / oracle_synth_bases so the two generators can be cross-checked.
//   scalars[i] = Fr::from_u512(w_0..w_7), w_j = splitmix64(seed * 0x100000001b3 + 8 i + j)
//                (from_u512 = d0*R^2 + d1*R^3: reference arithmetic/curves/src/derive/field.rs:29-48; this is how
//                 Fr::random draws a uniform scalar, bn256/fr.rs:159-170)
//   bases[i]   = [s0 + i d] G with (s0, d) = scalars 0 and 1 of stream `seed`, G = (1, 2) (bn256/curve.rs:66-67) —
//                distinct, non-identity curve points; each thread jumps to its run's first point with one
//                double-and-add, walks RUN mixed additions and normalises its run with one shared inversion
//                (the batch_normalize trick, derive/curve.rs:362-397). This is also the device-side shape of SRS
//                generation (poly/kzg/commitment.rs:209-233), a "next" row of SURVEY.md §8(f).
#include "internal.h"

namespace cqb {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

__device__ __forceinline__ Fr synth_scalar(uint64_t seed, uint64_t i) {
    Fr d0, d1;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint64_t a = splitmix64(seed * 0x100000001b3ULL + 8 * i + (uint64_t)j);
        uint64_t b = splitmix64(seed * 0x100000001b3ULL + 8 * i + (uint64_t)(j + 4));
        d0.l[2 * j] = (uint32_t)a; d0.l[2 * j + 1] = (uint32_t)(a >> 32);
        d1.l[2 * j] = (uint32_t)b; d1.l[2 * j + 1] = (uint32_t)(b >> 32);
    }
    return fp_add<FrP>(fp_mul<FrP>(d0, Fr::r2()), fp_mul<FrP>(d1, Fr::r3()));
}

__device__ __forceinline__ void st_fe(uint4* p, size_t i, const uint32_t* l) {
    p[2 * i] = make_uint4(l[0], l[1], l[2], l[3]);
    p[2 * i + 1] = make_uint4(l[4], l[5], l[6], l[7]);
}
__device__ __forceinline__ void ld_fe(const uint4* p, size_t i, uint32_t* l) {
    uint4 a = p[2 * i], b = p[2 * i + 1];
    l[0] = a.x; l[1] = a.y; l[2] = a.z; l[3] = a.w; l[4] = b.x; l[5] = b.y; l[6] = b.z; l[7] = b.w;
}

__global__ void synth_scalars_kernel(uint64_t seed, size_t start, size_t n, uint4* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = synth_scalar(seed, start + i);
    st_fe(out, i, s.l);
}

// [k]G for a Montgomery-form scalar k, G = (1,2)
__device__ G1Xyzz g1_mul_generator(const Fr& k_mont) {
    Fr k = fp_from_mont<FrP>(k_mont);
    Fq gx = Fq::one(), gy = fp_dbl<FqP>(Fq::one());
    G1Xyzz acc = G1Xyzz::identity();
    for (int i = 7; i >= 0; i--)
        for (int b = 31; b >= 0; b--) {
            acc = g1_double(acc);
            if ((k.l[i] >> b) & 1u) g1_madd(acc, gx, gy);
        }
    return acc;
}

// step[0..1] = affine [d]G
__global__ void synth_step_kernel(uint64_t seed, uint4* step) {
    if (threadIdx.x || blockIdx.x) return;
    G1Xyzz p = g1_mul_generator(synth_scalar(seed, 1));
    G1Affine a = g1_to_affine(p);
    st_fe(step, 0, a.x.l);
    st_fe(step, 1, a.y.l);
}

constexpr int GEN_RUN = 128;

// tmp layout per point: 4 Fq of XYZZ (128 B) then 1 Fq prefix product (32 B) = 10 uint4
__global__ void __launch_bounds__(128) synth_bases_kernel(uint64_t seed, size_t start, size_t n, const uint4* __restrict__ step,
                                                          uint4* __restrict__ tmp, uint4* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t p0 = t * GEN_RUN;
    if (p0 >= n) return;
    size_t cnt = (n - p0 < (size_t)GEN_RUN) ? (n - p0) : (size_t)GEN_RUN;
    // k = s0 + (start + p0) * d
    Fr s0 = synth_scalar(seed, 0), d = synth_scalar(seed, 1);
    uint64_t g0 = (uint64_t)(start + p0);
    Fr gi = Fr::zero();
    gi.l[0] = (uint32_t)g0; gi.l[1] = (uint32_t)(g0 >> 32);
    Fr k = fp_add<FrP>(s0, fp_mul<FrP>(d, fp_to_mont<FrP>(gi)));
    G1Xyzz P = g1_mul_generator(k);
    Fq dx, dy;
    ld_fe(step, 0, dx.l);
    ld_fe(step, 1, dy.l);
    Fq prod = Fq::one();
    for (size_t j = 0; j < cnt; j++) {
        uint4* slot = tmp + (p0 + j) * 10;
        st_fe(slot, 0, P.x.l); st_fe(slot, 1, P.y.l); st_fe(slot, 2, P.zz.l); st_fe(slot, 3, P.zzz.l);
        st_fe(slot, 4, prod.l);  // product of the z-factors of the points before this one
        if (!P.is_identity()) prod = fp_mul<FqP>(prod, fp_mul<FqP>(P.zz, P.zzz));
        g1_madd(P, dx, dy);
    }
    Fq inv = fp_inv<FqP>(prod);
    for (size_t j = cnt; j-- > 0;) {
        const uint4* slot = tmp + (p0 + j) * 10;
        Fq x, y, zz, zzz, pre;
        ld_fe(slot, 0, x.l); ld_fe(slot, 1, y.l); ld_fe(slot, 2, zz.l); ld_fe(slot, 3, zzz.l); ld_fe(slot, 4, pre.l);
        Fq ax = Fq::zero(), ay = Fq::zero();
        if (!zz.is_zero()) {
            Fq zi = fp_mul<FqP>(inv, pre);              // 1 / (zz * zzz) of this point
            inv = fp_mul<FqP>(inv, fp_mul<FqP>(zz, zzz));
            ax = fp_mul<FqP>(x, fp_mul<FqP>(zi, zzz));  // X / ZZ
            ay = fp_mul<FqP>(y, fp_mul<FqP>(zi, zz));   // Y / ZZZ
        }
        st_fe(out, 2 * (p0 + j), ax.l);
        st_fe(out, 2 * (p0 + j) + 1, ay.l);
    }
}

static Scratch g_gen_tmp;
void gen_release_all() { g_gen_tmp.release(); }

int synth_scalars_run(uint64_t seed, size_t start, size_t n, void* d_out) {
    if (n == 0) return 0;
    synth_scalars_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx().stream>>>(seed, start, n, (uint4*)d_out);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

int synth_bases_run(uint64_t seed, size_t start, size_t n, void* d_out) {
    if (n == 0) return 0;
    cudaStream_t st = ctx().stream;
    const size_t CHUNK = (size_t)1 << 22;
    size_t tmp_pts = n < CHUNK ? n : CHUNK;
    CQB_TRY(g_gen_tmp.ensure(64 + tmp_pts * 160));
    uint4* step = g_gen_tmp.as<uint4>();
    uint4* tmp = step + 4;
    synth_step_kernel<<<1, 32, 0, st>>>(seed, step);
    CQB_LAUNCHED();
    for (size_t off = 0; off < n; off += CHUNK) {
        size_t m = (n - off < CHUNK) ? (n - off) : CHUNK;
        size_t threads = (m + GEN_RUN - 1) / GEN_RUN;
        synth_bases_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(seed, start + off, m, step, tmp,
                                                                               (uint4*)d_out + off * 4);
        CQB_LAUNCHED();
    }
    CQB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cqb

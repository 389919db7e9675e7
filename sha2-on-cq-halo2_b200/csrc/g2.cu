// g2.cu — the G2 side of the SRS and of the CQ table commitment (keygen-time pieces of the path):
//   * [s]G2 of ParamsKZG (reference halo2_proofs/src/poly/kzg/commitment.rs:265-266) and the G2 powers [s^i]G2 of the table SRS
//     (:94-104, normalised :114-141);
//   * StaticTable::commit (plonk/static_lookup.rs:127-160): t = best_multiexp::<G2Affine>(table coefficients, srs_g2) (:146) and
//     zv = [s^N]G2 - G2 (:137) — MSMs of table size over G2, once per table.
// bn256 G2 is y^2 = x^3 + 3/(9+u) over Fq2 = Fq[u]/(u^2+1) (arithmetic/curves/src/bn256/fq2.rs, bn256/curve.rs:36-48, 85-129); the
// point formulas are the reference's generic Jacobian ones (derive/curve.rs:422-447, 853-893, 809-851). Only affine normal forms
// leave this file, so results are the reference's. Sizes are small (2^16) and this runs once per table: one thread per scalar
// multiplication (double-and-add over the canonical scalar's bits), a block tree for the sum, one shared multiplier copy per kernel
// to keep the code compact — no windowing.
#include "internal.h"

namespace cqb {

struct Fq2 { Fq c0, c1; };
struct G2Jac { Fq2 x, y, z; };

__device__ __noinline__ Fq g2_fmul(Fq a, Fq b) { return fp_mul<FqP>(a, b); }

__device__ __forceinline__ Fq2 fq2_zero() { Fq2 r; r.c0 = Fq::zero(); r.c1 = Fq::zero(); return r; }
__device__ __forceinline__ Fq2 fq2_one() { Fq2 r; r.c0 = Fq::one(); r.c1 = Fq::zero(); return r; }
__device__ __forceinline__ bool fq2_is_zero(const Fq2& a) { return a.c0.is_zero() && a.c1.is_zero(); }
__device__ __forceinline__ bool fq2_eq(const Fq2& a, const Fq2& b) { return a.c0 == b.c0 && a.c1 == b.c1; }
__device__ __forceinline__ Fq2 fq2_add(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = fp_add<FqP>(a.c0, b.c0); r.c1 = fp_add<FqP>(a.c1, b.c1); return r; }
__device__ __forceinline__ Fq2 fq2_sub(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = fp_sub<FqP>(a.c0, b.c0); r.c1 = fp_sub<FqP>(a.c1, b.c1); return r; }
__device__ __forceinline__ Fq2 fq2_dbl(const Fq2& a) { return fq2_add(a, a); }
__device__ __forceinline__ Fq2 fq2_neg(const Fq2& a) { Fq2 r; r.c0 = fp_neg<FqP>(a.c0); r.c1 = fp_neg<FqP>(a.c1); return r; }
// fq2.rs:161-170 (Karatsuba: three base-field multiplications)
__device__ __forceinline__ Fq2 fq2_mul(const Fq2& a, const Fq2& b) {
    Fq t1 = g2_fmul(a.c0, b.c0), t2 = g2_fmul(a.c1, b.c1);
    Fq t0 = g2_fmul(fp_add<FqP>(a.c0, a.c1), fp_add<FqP>(b.c0, b.c1));
    Fq2 r;
    r.c0 = fp_sub<FqP>(t1, t2);
    r.c1 = fp_sub<FqP>(t0, fp_add<FqP>(t1, t2));
    return r;
}
// fq2.rs:172-181 (two multiplications)
__device__ __forceinline__ Fq2 fq2_sqr(const Fq2& a) {
    Fq ab = g2_fmul(a.c0, a.c1);
    Fq c0 = g2_fmul(fp_sub<FqP>(a.c0, a.c1), fp_add<FqP>(a.c0, a.c1));
    Fq2 r;
    r.c0 = c0;  // (c0 - c1)(c0 + c1) - ab + ab
    r.c1 = fp_dbl<FqP>(ab);
    return r;
}
// fq2.rs:290-307: (c0 - c1 u) / (c0^2 + c1^2); 0 for 0
__device__ __forceinline__ Fq2 fq2_inv(const Fq2& a) {
    Fq t = fp_add<FqP>(g2_fmul(a.c0, a.c0), g2_fmul(a.c1, a.c1));
    t = fp_inv_safegcd<FqP>(t);
    Fq2 r;
    r.c0 = g2_fmul(a.c0, t);
    r.c1 = fp_neg<FqP>(g2_fmul(a.c1, t));
    return r;
}

__device__ __forceinline__ Fq fq_raw_to_mont(uint64_t l0, uint64_t l1, uint64_t l2, uint64_t l3) {
    Fq r;
    r.l[0] = (uint32_t)l0; r.l[1] = (uint32_t)(l0 >> 32); r.l[2] = (uint32_t)l1; r.l[3] = (uint32_t)(l1 >> 32);
    r.l[4] = (uint32_t)l2; r.l[5] = (uint32_t)(l2 >> 32); r.l[6] = (uint32_t)l3; r.l[7] = (uint32_t)(l3 >> 32);
    return fp_to_mont<FqP>(r);
}
// bn256/curve.rs:100-129
__device__ __forceinline__ void g2_generator(Fq2& x, Fq2& y) {
    x.c0 = fq_raw_to_mont(0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL);
    x.c1 = fq_raw_to_mont(0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL);
    y.c0 = fq_raw_to_mont(0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL);
    y.c1 = fq_raw_to_mont(0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL);
}

__device__ __forceinline__ G2Jac g2_identity() { G2Jac p; p.x = fq2_zero(); p.y = fq2_zero(); p.z = fq2_zero(); return p; }
// derive/curve.rs:422-447
__device__ __noinline__ G2Jac g2_double(G2Jac p) {
    if (fq2_is_zero(p.z)) return g2_identity();
    Fq2 a = fq2_sqr(p.x), b = fq2_sqr(p.y), c = fq2_sqr(b);
    Fq2 d = fq2_dbl(fq2_sub(fq2_sub(fq2_sqr(fq2_add(p.x, b)), a), c));
    Fq2 e = fq2_add(fq2_dbl(a), a);
    Fq2 f = fq2_sqr(e);
    G2Jac r;
    r.z = fq2_dbl(fq2_mul(p.z, p.y));
    r.x = fq2_sub(f, fq2_dbl(d));
    c = fq2_dbl(fq2_dbl(fq2_dbl(c)));
    r.y = fq2_sub(fq2_mul(e, fq2_sub(d, r.x)), c);
    return r;
}
// derive/curve.rs:809-851
__device__ __noinline__ G2Jac g2_add(G2Jac s, G2Jac o) {
    if (fq2_is_zero(s.z)) return o;
    if (fq2_is_zero(o.z)) return s;
    Fq2 z1z1 = fq2_sqr(s.z), z2z2 = fq2_sqr(o.z);
    Fq2 u1 = fq2_mul(s.x, z2z2), u2 = fq2_mul(o.x, z1z1);
    Fq2 s1 = fq2_mul(fq2_mul(s.y, z2z2), o.z), s2 = fq2_mul(fq2_mul(o.y, z1z1), s.z);
    if (fq2_eq(u1, u2)) {
        if (fq2_eq(s1, s2)) return g2_double(s);
        return g2_identity();
    }
    Fq2 h = fq2_sub(u2, u1);
    Fq2 i = fq2_sqr(fq2_dbl(h));
    Fq2 j = fq2_mul(h, i);
    Fq2 r = fq2_dbl(fq2_sub(s2, s1));
    Fq2 v = fq2_mul(u1, i);
    G2Jac q;
    q.x = fq2_sub(fq2_sub(fq2_sub(fq2_sqr(r), j), v), v);
    q.y = fq2_sub(fq2_mul(r, fq2_sub(v, q.x)), fq2_dbl(fq2_mul(s1, j)));
    q.z = fq2_mul(fq2_sub(fq2_sub(fq2_sqr(fq2_add(s.z, o.z)), z1z1), z2z2), h);
    return q;
}
// derive/curve.rs:853-893 (the affine operand is not the identity)
__device__ __noinline__ G2Jac g2_madd(G2Jac s, Fq2 x2, Fq2 y2) {
    if (fq2_is_zero(s.z)) { G2Jac r; r.x = x2; r.y = y2; r.z = fq2_one(); return r; }
    Fq2 z1z1 = fq2_sqr(s.z);
    Fq2 u2 = fq2_mul(x2, z1z1);
    Fq2 s2 = fq2_mul(fq2_mul(y2, z1z1), s.z);
    if (fq2_eq(s.x, u2)) {
        if (fq2_eq(s.y, s2)) return g2_double(s);
        return g2_identity();
    }
    Fq2 h = fq2_sub(u2, s.x);
    Fq2 hh = fq2_sqr(h);
    Fq2 i = fq2_dbl(fq2_dbl(hh));
    Fq2 j = fq2_mul(h, i);
    Fq2 r = fq2_dbl(fq2_sub(s2, s.y));
    Fq2 v = fq2_mul(s.x, i);
    G2Jac q;
    q.x = fq2_sub(fq2_sub(fq2_sub(fq2_sqr(r), j), v), v);
    q.y = fq2_sub(fq2_mul(r, fq2_sub(v, q.x)), fq2_dbl(fq2_mul(s.y, j)));
    q.z = fq2_sub(fq2_sub(fq2_sqr(fq2_add(s.z, h)), z1z1), hh);
    return q;
}

__device__ __forceinline__ Fq ld_fq_g(const uint4* p) {
    uint4 a = p[0], b = p[1];
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fq_g(uint4* p, const Fq& v) {
    p[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ Fq2 ld_fq2(const uint4* p) { Fq2 r; r.c0 = ld_fq_g(p); r.c1 = ld_fq_g(p + 2); return r; }
__device__ __forceinline__ void st_fq2(uint4* p, const Fq2& v) { st_fq_g(p, v.c0); st_fq_g(p + 2, v.c1); }
__device__ __forceinline__ G2Jac ld_g2j(const uint4* p) { G2Jac r; r.x = ld_fq2(p); r.y = ld_fq2(p + 4); r.z = ld_fq2(p + 8); return r; }
__device__ __forceinline__ void st_g2j(uint4* p, const G2Jac& v) { st_fq2(p, v.x); st_fq2(p + 4, v.y); st_fq2(p + 8, v.z); }

// out[i] = [scalars[i]] bases[i] (bases == nullptr: the generator), Jacobian. derive/curve.rs:1019-1040: bits of to_repr(), MSB first.
__global__ void __launch_bounds__(64) g2_scalar_mul_kernel(const uint4* __restrict__ bases, const uint4* __restrict__ scalars, size_t n,
                                                           uint4* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fq2 bx, by;
    if (bases) { bx = ld_fq2(bases + i * 8); by = ld_fq2(bases + i * 8 + 4); }
    else g2_generator(bx, by);
    Fr k;
    {
        uint4 a = scalars[2 * i], b = scalars[2 * i + 1];
        k.l[0] = a.x; k.l[1] = a.y; k.l[2] = a.z; k.l[3] = a.w; k.l[4] = b.x; k.l[5] = b.y; k.l[6] = b.z; k.l[7] = b.w;
    }
    k = fp_from_mont<FrP>(k);
    G2Jac acc = g2_identity();
    if (!(fq2_is_zero(bx) && fq2_is_zero(by))) {
#pragma unroll 1
        for (int bit = 253; bit >= 0; bit--) {
            acc = g2_double(acc);
            if ((k.l[bit >> 5] >> (bit & 31)) & 1u) acc = g2_madd(acc, bx, by);
        }
    }
    st_g2j(out + i * 12, acc);
}

// CTA j adds inputs [j * 2048, (j + 1) * 2048) into out[j]
__global__ void __launch_bounds__(64) g2_sum_kernel(const uint4* __restrict__ in, size_t count, uint4* __restrict__ out) {
    __shared__ uint4 sm[64 * 12];
    const size_t lo = (size_t)blockIdx.x * 2048, hi = lo + 2048 < count ? lo + 2048 : count;
    G2Jac acc = g2_identity();
    for (size_t k = lo + threadIdx.x; k < hi; k += 64) acc = g2_add(acc, ld_g2j(in + k * 12));
    st_g2j(sm + threadIdx.x * 12, acc);
    __syncthreads();
#pragma unroll 1
    for (int half = 32; half >= 1; half >>= 1) {
        if ((int)threadIdx.x < half) {
            G2Jac a = ld_g2j(sm + threadIdx.x * 12), b = ld_g2j(sm + (threadIdx.x + half) * 12);
            st_g2j(sm + threadIdx.x * 12, g2_add(a, b));
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) st_g2j(out + (size_t)blockIdx.x * 12, ld_g2j(sm));
}

// derive/curve.rs:399-412 to_affine per point (identity -> zeros); 128-byte affine out, then (count == 1 only) an identity flag word
__global__ void __launch_bounds__(64) g2_normalize_kernel(const uint4* __restrict__ in, size_t n, uint4* __restrict__ out, int flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G2Jac p = ld_g2j(in + i * 12);
    Fq2 ax = fq2_zero(), ay = fq2_zero();
    const bool inf = fq2_is_zero(p.z);
    if (!inf) {
        Fq2 zi = fq2_inv(p.z), zi2 = fq2_sqr(zi);
        ax = fq2_mul(p.x, zi2);
        ay = fq2_mul(p.y, fq2_mul(zi2, zi));
    }
    st_fq2(out + i * 8, ax);
    st_fq2(out + i * 8 + 4, ay);
    if (flag) out[8] = make_uint4(inf ? 1u : 0u, 0, 0, 0);
}

static Scratch g_g2_tmp;
void g2_release_all() { g_g2_tmp.release(); }

// d_out_affine[i] = [scalars[i]] bases[i] (d_bases == nullptr: [scalars[i]] G2), affine
int g2_mul_run(const void* d_bases, const void* d_scalars, size_t n, void* d_out_affine) {
    if (n == 0) return 0;
    cudaStream_t st = ctx().stream;
    CQB_TRY(g_g2_tmp.ensure(n * 192));
    g2_scalar_mul_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>((const uint4*)d_bases, (const uint4*)d_scalars, n, g_g2_tmp.as<uint4>());
    CQB_LAUNCHED();
    g2_normalize_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(g_g2_tmp.as<uint4>(), n, (uint4*)d_out_affine, 0);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

// d_out: 128-byte affine sum_i [scalars[i]] bases[i], then a uint32 identity flag (144 bytes)
int g2_msm_run(const void* d_bases, const void* d_scalars, size_t n, void* d_out) {
    cudaStream_t st = ctx().stream;
    const size_t lvl1 = (n + 2047) / 2048;
    CQB_TRY(g_g2_tmp.ensure((n + lvl1 + 2) * 192 + 192));
    uint4* pts = g_g2_tmp.as<uint4>();
    if (n == 0) {
        CQB_CUDA(cudaMemsetAsync(pts, 0, 192, st));
    } else {
        g2_scalar_mul_kernel<<<(unsigned)((n + 63) / 64), 64, 0, st>>>((const uint4*)d_bases, (const uint4*)d_scalars, n, pts);
        CQB_LAUNCHED();
    }
    size_t count = n ? n : 1;
    uint4* cur = pts;
    uint4* nxt = pts + (n + 1) * 12;
    while (count > 1) {
        const size_t blocks = (count + 2047) / 2048;
        g2_sum_kernel<<<(unsigned)blocks, 64, 0, st>>>(cur, count, nxt);
        CQB_LAUNCHED();
        uint4* t = cur; cur = nxt; nxt = t;  // ping-pong: the next level is at most count / 2048 + 1 points and fits either region
        count = blocks;
    }
    g2_normalize_kernel<<<1, 64, 0, st>>>(cur, 1, (uint4*)d_out, 1);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cqb

// fp.cuh — 8x32-bit-limb Montgomery arithmetic for the two bn256 prime fields (Fr scalar field, Fq base field).
//
// Semantics contract (results are limb-identical to the reference because every value is kept fully reduced in
// [0, p) and Montgomery multiplication has a unique canonical answer a*b*2^-256 mod p):
//   add/sub/neg/double      -> reference arithmetic/curves/src/derive/field.rs:351-430, 488-500
//   mul / square            -> reference derive/field.rs:502-562 (sparse CIOS) and :358-393 (square + montgomery_reduce :564-617)
//   from_mont (to canonical)-> reference derive/field.rs:432-470 (montgomery_reduce_short), used by to_repr bn256/fr.rs:241-257
// Constants                 -> reference bn256/fr.rs:29-118, bn256/fq.rs:28-90 (64-bit limbs split into 32-bit halves;
//                              INV32 = INV mod 2^32).
//
// This is NOT a translation of the reference's 4x64 code: the limb width (32), the even/odd split carry chains and the
// PTX mad.lo.cc/madc.hi.cc pairing (which ptxas fuses into IMAD.WIDE on sm_100a) are chosen for the B200 integer pipe.
//
// The same source compiles for the host (carry flag emulated in a thread-local) so the exact limb algorithm can be
// unit-tested on a box without a GPU (tests/test_fp_host.py builds tools/fp_host_check.cu with nvcc -x cu host path).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CQB_HD __host__ __device__ __forceinline__
#define CQB_D __device__ __forceinline__
#else
#define CQB_HD inline
#define CQB_D inline
#endif

namespace cqb {

// ---------------------------------------------------------------------------------------------------------------------
// carry-chain primitives: PTX on the device, emulation on the host
// ---------------------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define CQB_ASM_CC 1
#else
#define CQB_ASM_CC 0
static thread_local uint32_t cqb_cf = 0;  // emulated carry/borrow flag (host only)
#endif

CQB_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
CQB_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#if CQB_ASM_CC
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

#if CQB_ASM_CC
#define CQB_OP3(name, ptx)                                                                  \
    CQB_D uint32_t name(uint32_t a, uint32_t b) {                                           \
        uint32_t r;                                                                         \
        asm volatile(ptx " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                        \
        return r;                                                                           \
    }
#define CQB_OP4(name, ptx)                                                                  \
    CQB_D uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                               \
        uint32_t r;                                                                         \
        asm volatile(ptx " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));            \
        return r;                                                                           \
    }
CQB_OP3(add_cc, "add.cc.u32")
CQB_OP3(addc_cc, "addc.cc.u32")
CQB_OP3(addc, "addc.u32")
CQB_OP3(sub_cc, "sub.cc.u32")
CQB_OP3(subc_cc, "subc.cc.u32")
CQB_OP3(subc, "subc.u32")
CQB_OP4(mad_lo_cc, "mad.lo.cc.u32")
CQB_OP4(madc_lo_cc, "madc.lo.cc.u32")
CQB_OP4(mad_hi_cc, "mad.hi.cc.u32")
CQB_OP4(madc_hi_cc, "madc.hi.cc.u32")
CQB_OP4(madc_hi, "madc.hi.u32")
CQB_OP4(madc_lo, "madc.lo.u32")
#undef CQB_OP3
#undef CQB_OP4
#else
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; cqb_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + cqb_cf; cqb_cf = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cqb_cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; cqb_cf = (uint32_t)(s >> 63); return (uint32_t)s; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - cqb_cf; cqb_cf = (uint32_t)(s >> 63); return (uint32_t)s; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cqb_cf; }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(a * b, c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(a * b, c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return addc(a * b, c); }
#endif

// ---------------------------------------------------------------------------------------------------------------------
// field parameters. mod(i)/r(i)/r2(i) are pure expressions of i so that, after full unrolling, they fold to immediates.
// ---------------------------------------------------------------------------------------------------------------------
#define CQB_SEL8(i, a0, a1, a2, a3, a4, a5, a6, a7) \
    ((i) == 0 ? a0 : (i) == 1 ? a1 : (i) == 2 ? a2 : (i) == 3 ? a3 : (i) == 4 ? a4 : (i) == 5 ? a5 : (i) == 6 ? a6 : a7)

// scalar field r (reference bn256/fr.rs:29-66)
struct FrP {
    static CQB_HD constexpr uint32_t mod(int i) {
        return CQB_SEL8(i, 0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u);
    }
    static CQB_HD constexpr uint32_t inv() { return 0xefffffffu; }  // INV mod 2^32, fr.rs:39
    static CQB_HD constexpr uint32_t r(int i) {                      // R = 2^256 mod r, fr.rs:43-48
        return CQB_SEL8(i, 0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u);
    }
    static CQB_HD constexpr uint32_t r2(int i) {                     // R^2, fr.rs:52-57
        return CQB_SEL8(i, 0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u);
    }
    static CQB_HD constexpr uint32_t r3(int i) {                     // R^3, fr.rs:61-66
        return CQB_SEL8(i, 0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu);
    }
};

// base field q (reference bn256/fq.rs:28-60)
struct FqP {
    static CQB_HD constexpr uint32_t mod(int i) {
        return CQB_SEL8(i, 0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u);
    }
    static CQB_HD constexpr uint32_t inv() { return 0xe4866389u; }  // INV mod 2^32, fq.rs:37
    static CQB_HD constexpr uint32_t r(int i) {                      // fq.rs:40-45
        return CQB_SEL8(i, 0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u);
    }
    static CQB_HD constexpr uint32_t r2(int i) {                     // fq.rs:48-53
        return CQB_SEL8(i, 0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u);
    }
    static CQB_HD constexpr uint32_t r3(int i) {                     // fq.rs:56-61
        return CQB_SEL8(i, 0xda1530dfu, 0xb1cd6dafu, 0xa7283db6u, 0x62f210e6u, 0x0ada0afbu, 0xef7f0b0cu, 0x2d592544u, 0x20fd6e90u);
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// field element: 8 little-endian 32-bit limbs == the reference's [u64;4] reinterpret-cast on a little-endian host
// ---------------------------------------------------------------------------------------------------------------------
template <class P>
struct Fp {
    uint32_t l[8];

    static CQB_HD Fp zero() {
        Fp z;
#pragma unroll
        for (int i = 0; i < 8; i++) z.l[i] = 0;
        return z;
    }
    static CQB_HD Fp one() {  // Montgomery one = R
        Fp z;
#pragma unroll
        for (int i = 0; i < 8; i++) z.l[i] = P::r(i);
        return z;
    }
    static CQB_HD Fp r2() {
        Fp z;
#pragma unroll
        for (int i = 0; i < 8; i++) z.l[i] = P::r2(i);
        return z;
    }
    static CQB_HD Fp r3() {
        Fp z;
#pragma unroll
        for (int i = 0; i < 8; i++) z.l[i] = P::r3(i);
        return z;
    }
    CQB_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= l[i];
        return o == 0;
    }
    CQB_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) o |= (l[i] ^ b.l[i]);
        return o == 0;
    }
    CQB_HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// r = a - p if a >= p else a      (a < 2p)
template <class P>
CQB_HD void fp_reduce_once(uint32_t* a) {
    uint32_t t[8];
    t[0] = sub_cc(a[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = subc_cc(a[i], P::mod(i));
    uint32_t borrow = subc(0u, 0u);  // 0xffffffff if a < p
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = borrow ? a[i] : t[i];
}

// reference derive/field.rs:488-500 (sparse add: top limb cannot overflow because p < 2^254)
template <class P>
CQB_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
    r.l[7] = addc(a.l[7], b.l[7]);
    fp_reduce_once<P>(r.l);
    return r;
}

// reference derive/field.rs:395-412
template <class P>
CQB_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = subc(0u, 0u);  // all-ones mask when a < b
    r.l[0] = add_cc(r.l[0], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; i++) r.l[i] = addc_cc(r.l[i], P::mod(i) & borrow);
    r.l[7] = addc(r.l[7], P::mod(7) & borrow);
    return r;
}

// reference derive/field.rs:351-354
template <class P>
CQB_HD Fp<P> fp_dbl(const Fp<P>& a) { return fp_add<P>(a, a); }

// reference derive/field.rs:414-430
template <class P>
CQB_HD Fp<P> fp_neg(const Fp<P>& a) {
    Fp<P> r;
    r.l[0] = sub_cc(P::mod(0), a.l[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) r.l[i] = subc_cc(P::mod(i), a.l[i]);
    r.l[7] = subc(P::mod(7), a.l[7]);
    uint32_t nz = a.is_zero() ? 0u : 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] &= nz;
    return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// Montgomery multiplication, CIOS over 32-bit limbs with two interleaved accumulators.
//
// Products x[j]*y are 64 bits wide and land on limb j. Products of even j are chained into the accumulator aligned
// on limb 0 (lo,hi,lo,hi,... = one unbroken carry chain); products of odd j go to a second accumulator aligned on limb 1.
// After the reduction step limb 0 is zero and the frame moves up one limb: the two accumulators swap roles and the one
// that becomes limb-1 aligned is re-read two limbs higher, so no data is ever moved. Every (mad.lo.cc, madc.hi.cc) pair
// shares its operands, which is what lets ptxas emit one IMAD.WIDE per 32x32->64 product.
// ---------------------------------------------------------------------------------------------------------------------

// acc[0..7] = x[0,2,4,6]*y (fresh, no carries), acc[8] = 0
CQB_HD void row_mul(uint32_t* acc, const uint32_t* x, uint32_t y) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        acc[j] = mul_lo(x[j], y);
        acc[j + 1] = mul_hi(x[j], y);
    }
    acc[8] = 0;
}

// acc[0..8] += x[0,2,4,6]*y  (fresh carry chain)
CQB_HD void row_mad(uint32_t* acc, const uint32_t* x, uint32_t y) {
    acc[0] = mad_lo_cc(x[0], y, acc[0]);
    acc[1] = madc_hi_cc(x[0], y, acc[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        acc[j] = madc_lo_cc(x[j], y, acc[j]);
        acc[j + 1] = madc_hi_cc(x[j], y, acc[j + 1]);
    }
    acc[8] = addc(acc[8], 0u);
}

// acc[0..8] = (acc >> 64) + x[0,2,4,6]*y, with the incoming carry flag consumed at limb 0
CQB_HD void row_madc_shift2(uint32_t* acc, const uint32_t* x, uint32_t y) {
#pragma unroll
    for (int j = 0; j < 6; j += 2) {
        acc[j] = madc_lo_cc(x[j], y, acc[j + 2]);
        acc[j + 1] = madc_hi_cc(x[j], y, acc[j + 3]);
    }
    acc[6] = madc_lo_cc(x[6], y, acc[8]);
    acc[7] = madc_hi_cc(x[6], y, 0u);
    acc[8] = addc(0u, 0u);
}

template <class P>
CQB_HD void mod_limbs(uint32_t* m) {
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::mod(i);
}

// reference derive/field.rs:502-562: result = a*b*2^-256 mod p, fully reduced
template <class P>
CQB_HD Fp<P> fp_mul_kar(const Fp<P>& a, const Fp<P>& b);
template <class P>
CQB_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
#ifdef CQB_KARATSUBA
    return fp_mul_kar<P>(a, b);
#else
    uint32_t ev[9], od[9], p[8];
    mod_limbs<P>(p);
    uint32_t m;

    // i = 0 : ev is limb-0 aligned, od is limb-1 aligned
    row_mul(ev, a.l, b.l[0]);
    row_mul(od, a.l + 1, b.l[0]);
    m = ev[0] * P::inv();
    row_mad(od, p + 1, m);
    row_mad(ev, p, m);

#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        // odd i: od becomes limb-0 aligned, ev (read two limbs up) limb-1 aligned; ev[1] is the orphan at limb 0
        od[0] = add_cc(od[0], ev[1]);
        row_madc_shift2(ev, a.l + 1, b.l[i]);
        row_mad(od, a.l, b.l[i]);
        m = od[0] * P::inv();
        row_mad(ev, p + 1, m);
        row_mad(od, p, m);
        if (i + 1 < 8) {
            // even i+1: roles swap back
            ev[0] = add_cc(ev[0], od[1]);
            row_madc_shift2(od, a.l + 1, b.l[i + 1]);
            row_mad(ev, a.l, b.l[i + 1]);
            m = ev[0] * P::inv();
            row_mad(od, p + 1, m);
            row_mad(ev, p, m);
        }
    }
    // after i = 7: od limb-0 aligned with od[0] == 0, ev limb-1 aligned. result = ev + (od >> 32)
    Fp<P> r;
    r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < 7; k++) r.l[k] = addc_cc(ev[k], od[k + 1]);
    r.l[7] = addc(ev[7], od[8]);
    fp_reduce_once<P>(r.l);
    return r;
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant: separated product + reduction with one level of Karatsuba on the 8x8-limb product (3 x 16 = 48 wide products
// instead of 64; the reduction keeps 64 + 8). The multiplier ("fmaheavy") pipe is the bottleneck of every kernel here
// while the ALU pipe idles at ~30 %, so trading 16 IMAD.WIDE for ~100 IADD3 is a net win if ptxas keeps the chains fused.
// Same canonical result as fp_mul.
// ---------------------------------------------------------------------------------------------------------------------
template <int N>
CQB_HD void rowN_mul(uint32_t* acc, const uint32_t* x, uint32_t y) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        const uint64_t pr = (uint64_t)x[j] * (uint64_t)y;  // mul.wide.u32 -> one IMAD.WIDE
        acc[j] = (uint32_t)pr;
        acc[j + 1] = (uint32_t)(pr >> 32);
    }
    acc[N] = 0;
}
template <int N>
CQB_HD void rowN_mad(uint32_t* acc, const uint32_t* x, uint32_t y) {
    acc[0] = mad_lo_cc(x[0], y, acc[0]);
    acc[1] = madc_hi_cc(x[0], y, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        acc[j] = madc_lo_cc(x[j], y, acc[j]);
        acc[j + 1] = madc_hi_cc(x[j], y, acc[j + 1]);
    }
    acc[N] = addc(acc[N], 0u);
}
template <int N>
CQB_HD void rowN_madc_shift2(uint32_t* acc, const uint32_t* x, uint32_t y) {
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        acc[j] = madc_lo_cc(x[j], y, acc[j + 2]);
        acc[j + 1] = madc_hi_cc(x[j], y, acc[j + 3]);
    }
    acc[N - 2] = madc_lo_cc(x[N - 2], y, acc[N]);
    acc[N - 1] = madc_hi_cc(x[N - 2], y, 0u);
    acc[N] = addc(0u, 0u);
}

// r[0..2N) = a[0..N) * b[0..N), N even: the even/odd two-accumulator scheme of fp_mul without the reduction rows; the
// limb that leaves the frame after each row is a finished limb of the product.
template <int N>
CQB_HD void mul_nxn(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t ev[N + 1], od[N + 1];
    rowN_mul<N>(ev, a, b[0]);
    rowN_mul<N>(od, a + 1, b[0]);
    r[0] = ev[0];
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        od[0] = add_cc(od[0], ev[1]);
        rowN_madc_shift2<N>(ev, a + 1, b[i]);
        rowN_mad<N>(od, a, b[i]);
        r[i] = od[0];
        if (i + 1 < N) {
            ev[0] = add_cc(ev[0], od[1]);
            rowN_madc_shift2<N>(od, a + 1, b[i + 1]);
            rowN_mad<N>(ev, a, b[i + 1]);
            r[i + 1] = ev[0];
        }
    }
    r[N] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r[N + k] = addc_cc(ev[k], od[k + 1]);
    r[2 * N - 1] = addc(ev[N - 1], od[N]);
}

// t[0..16) = a * b with one Karatsuba level over 4-limb halves
CQB_HD void mul_8x8_karatsuba(uint32_t* t, const uint32_t* a, const uint32_t* b) {
    uint32_t z0[8], z2[8], z1[9], sa[4], sb[4];
    mul_nxn<4>(z0, a, b);
    mul_nxn<4>(z2, a + 4, b + 4);
    sa[0] = add_cc(a[0], a[4]); sa[1] = addc_cc(a[1], a[5]); sa[2] = addc_cc(a[2], a[6]); sa[3] = addc_cc(a[3], a[7]);
    const uint32_t ca = addc(0u, 0u);
    sb[0] = add_cc(b[0], b[4]); sb[1] = addc_cc(b[1], b[5]); sb[2] = addc_cc(b[2], b[6]); sb[3] = addc_cc(b[3], b[7]);
    const uint32_t cb = addc(0u, 0u);
    mul_nxn<4>(z1, sa, sb);
    z1[8] = ca & cb;
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    // (sa + ca 2^128)(sb + cb 2^128) = sa sb + (ca sb + cb sa) 2^128 + ca cb 2^256
    z1[4] = add_cc(z1[4], sb[0] & ma); z1[5] = addc_cc(z1[5], sb[1] & ma); z1[6] = addc_cc(z1[6], sb[2] & ma);
    z1[7] = addc_cc(z1[7], sb[3] & ma); z1[8] = addc(z1[8], 0u);
    z1[4] = add_cc(z1[4], sa[0] & mb); z1[5] = addc_cc(z1[5], sa[1] & mb); z1[6] = addc_cc(z1[6], sa[2] & mb);
    z1[7] = addc_cc(z1[7], sa[3] & mb); z1[8] = addc(z1[8], 0u);
    // z1 -= z0 + z2  (the middle term a0 b1 + a1 b0 is non-negative and < 2^257)
    z1[0] = sub_cc(z1[0], z0[0]);
#pragma unroll
    for (int k = 1; k < 8; k++) z1[k] = subc_cc(z1[k], z0[k]);
    z1[8] = subc(z1[8], 0u);
    z1[0] = sub_cc(z1[0], z2[0]);
#pragma unroll
    for (int k = 1; k < 8; k++) z1[k] = subc_cc(z1[k], z2[k]);
    z1[8] = subc(z1[8], 0u);
    // t = z0 + z1 2^128 + z2 2^256
#pragma unroll
    for (int k = 0; k < 4; k++) t[k] = z0[k];
    t[4] = add_cc(z0[4], z1[0]); t[5] = addc_cc(z0[5], z1[1]); t[6] = addc_cc(z0[6], z1[2]); t[7] = addc_cc(z0[7], z1[3]);
    t[8] = addc_cc(z2[0], z1[4]); t[9] = addc_cc(z2[1], z1[5]); t[10] = addc_cc(z2[2], z1[6]); t[11] = addc_cc(z2[3], z1[7]);
    t[12] = addc_cc(z2[4], z1[8]); t[13] = addc_cc(z2[5], 0u); t[14] = addc_cc(z2[6], 0u); t[15] = addc(z2[7], 0u);
}

// Montgomery reduction of a 16-limb product t < p * 2^256: (t + M p) / 2^256 with M chosen limb by limb, then one
// conditional subtraction (reference derive/field.rs:564-617 montgomery_reduce, restructured into the even/odd chains)
template <class P>
CQB_HD Fp<P> fp_reduce16(const uint32_t* t) {
    uint32_t ev[9], od[9], p[8];
    mod_limbs<P>(p);
#pragma unroll
    for (int k = 0; k < 8; k++) ev[k] = t[k];
    ev[8] = 0;
    uint32_t m = ev[0] * P::inv();
    rowN_mul<8>(od, p + 1, m);
    rowN_mad<8>(ev, p, m);
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        od[0] = add_cc(od[0], ev[1]);
        m = od[0] * P::inv();  // mul.lo does not touch the carry flag consumed by the chain below
        rowN_madc_shift2<8>(ev, p + 1, m);
        rowN_mad<8>(od, p, m);
        if (i + 1 < 8) {
            ev[0] = add_cc(ev[0], od[1]);
            m = ev[0] * P::inv();
            rowN_madc_shift2<8>(od, p + 1, m);
            rowN_mad<8>(ev, p, m);
        }
    }
    Fp<P> r;
    // U = ev + (od >> 32) <= p ; result = U + t_hi < 2p
    r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < 7; k++) r.l[k] = addc_cc(ev[k], od[k + 1]);
    r.l[7] = addc(ev[7], od[8]);
    r.l[0] = add_cc(r.l[0], t[8]);
#pragma unroll
    for (int k = 1; k < 7; k++) r.l[k] = addc_cc(r.l[k], t[8 + k]);
    r.l[7] = addc(r.l[7], t[15]);
    fp_reduce_once<P>(r.l);
    return r;
}

template <class P>
CQB_HD Fp<P> fp_mul_kar(const Fp<P>& a, const Fp<P>& b) {
    uint32_t t[16];
    mul_8x8_karatsuba(t, a.l, b.l);
    return fp_reduce16<P>(t);
}


// t[0..16) = a^2 with 36 wide products instead of 64: the 28 off-diagonal products a_i a_j (i < j) are accumulated once,
// doubled with funnel shifts on the ALU pipe, and the 8 diagonal products are added by one carry chain. Product a_i a_j
// lands on limb i+j: those with i+j even go to the limb-0-aligned accumulator E, those with i+j odd to the limb-1-aligned
// accumulator O (O[k] is limb k+1), so that within a row consecutive j of equal parity form one unbroken lo,hi,lo,hi
// carry chain (one IMAD.WIDE per product, as in fp_mul). Rows are visited in increasing i, which makes the limb that
// receives a chain's final carry one that no product has been written to yet: a single addc, no carry propagation.
CQB_HD void sqr_8(uint32_t* t, const uint32_t* a) {
    uint32_t E[16], O[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
    for (int i = 0; i < 7; i++) {
        if (i + 2 < 8) {  // j = i+2, i+4, ... : limb i+j of E
            E[2 * i + 2] = mad_lo_cc(a[i], a[i + 2], E[2 * i + 2]);
            E[2 * i + 3] = madc_hi_cc(a[i], a[i + 2], E[2 * i + 3]);
            int end = 2 * i + 4;
#pragma unroll
            for (int j = i + 4; j < 8; j += 2) {
                E[i + j] = madc_lo_cc(a[i], a[j], E[i + j]);
                E[i + j + 1] = madc_hi_cc(a[i], a[j], E[i + j + 1]);
                end = i + j + 2;
            }
            E[end] = addc(E[end], 0u);
        }
        {  // j = i+1, i+3, ... : index i+j-1 of O
            O[2 * i] = mad_lo_cc(a[i], a[i + 1], O[2 * i]);
            O[2 * i + 1] = madc_hi_cc(a[i], a[i + 1], O[2 * i + 1]);
            int end = 2 * i + 2;
#pragma unroll
            for (int j = i + 3; j < 8; j += 2) {
                O[i + j - 1] = madc_lo_cc(a[i], a[j], O[i + j - 1]);
                O[i + j] = madc_hi_cc(a[i], a[j], O[i + j]);
                end = i + j + 1;
            }
            O[end] = addc(O[end], 0u);
        }
    }
    // S = E + (O << 32)  (E[0] = E[1] = 0), S < 2^511
    uint32_t s[16];
    s[0] = 0;
    s[1] = O[0];
    s[2] = add_cc(E[2], O[1]);
#pragma unroll
    for (int k = 3; k < 15; k++) s[k] = addc_cc(E[k], O[k - 1]);
    s[15] = addc(E[15], O[14]);
    // t = 2 S + sum_i a_i^2 2^(64 i)
    t[0] = mad_lo_cc(a[0], a[0], 0u);
    t[1] = madc_hi_cc(a[0], a[0], s[1] << 1);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        t[2 * i] = madc_lo_cc(a[i], a[i], (s[2 * i] << 1) | (s[2 * i - 1] >> 31));
        t[2 * i + 1] = madc_hi_cc(a[i], a[i], (s[2 * i + 1] << 1) | (s[2 * i] >> 31));
    }
}

// reference derive/field.rs:358-393 (square + montgomery_reduce): same canonical result as mul(a, a), 36 + 72 wide
// multiplies instead of 136.
template <class P>
CQB_HD Fp<P> fp_sqr(const Fp<P>& a) {
#ifdef CQB_NO_FAST_SQR
    return fp_mul<P>(a, a);
#else
    uint32_t t[16];
    sqr_8(t, a.l);
    return fp_reduce16<P>(t);
#endif
}

// a*b + c*d (Montgomery form, fully reduced) with ONE reduction: CIOS with two product rows per reduction row
// (2 x 64 + 72 wide multiplies instead of 2 x 136). Bound: a, b, c, d <= p < 2^254, so the running sum stays below 4p and
// the result before the conditional subtraction below (2 p^2 + 2^256 p) / 2^256 < 2p.
template <class P>
CQB_HD Fp<P> fp_mul2(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
    uint32_t ev[9], od[9], p[8];
    mod_limbs<P>(p);
    uint32_t m;
    row_mul(ev, a.l, b.l[0]);
    row_mul(od, a.l + 1, b.l[0]);
    row_mad(ev, c.l, d.l[0]);
    row_mad(od, c.l + 1, d.l[0]);
    m = ev[0] * P::inv();
    row_mad(od, p + 1, m);
    row_mad(ev, p, m);
#pragma unroll
    for (int i = 1; i < 8; i += 2) {
        od[0] = add_cc(od[0], ev[1]);
        row_madc_shift2(ev, a.l + 1, b.l[i]);
        row_mad(od, a.l, b.l[i]);
        row_mad(ev, c.l + 1, d.l[i]);
        row_mad(od, c.l, d.l[i]);
        m = od[0] * P::inv();
        row_mad(ev, p + 1, m);
        row_mad(od, p, m);
        if (i + 1 < 8) {
            ev[0] = add_cc(ev[0], od[1]);
            row_madc_shift2(od, a.l + 1, b.l[i + 1]);
            row_mad(ev, a.l, b.l[i + 1]);
            row_mad(od, c.l + 1, d.l[i + 1]);
            row_mad(ev, c.l, d.l[i + 1]);
            m = ev[0] * P::inv();
            row_mad(od, p + 1, m);
            row_mad(ev, p, m);
        }
    }
    Fp<P> r;
    r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < 7; k++) r.l[k] = addc_cc(ev[k], od[k + 1]);
    r.l[7] = addc(ev[7], od[8]);
    fp_reduce_once<P>(r.l);
    return r;
}

// Montgomery -> canonical integer (multiply by 1): reference derive/field.rs:432-470 montgomery_reduce_short
template <class P>
CQB_HD Fp<P> fp_from_mont(const Fp<P>& a) {
    Fp<P> one_raw = Fp<P>::zero();
    one_raw.l[0] = 1;
    return fp_mul<P>(a, one_raw);
}
// canonical integer (< p) -> Montgomery: multiply by R^2 (reference derive/field.rs:50-53 from_raw)
template <class P>
CQB_HD Fp<P> fp_to_mont(const Fp<P>& a) { return fp_mul<P>(a, Fp<P>::r2()); }

// a^e for a 256-bit exponent given as 8 limbs (square-and-multiply, MSB first); used for inversion a^(p-2)
// (reference bn256/fr.rs:200-209, fq.rs invert) — variable time is fine here, nothing is secret on this path.
template <class P>
CQB_HD Fp<P> fp_pow(const Fp<P>& a, const uint32_t* e) {
    Fp<P> r = Fp<P>::one();
    bool started = false;
    for (int i = 7; i >= 0; i--) {
        for (int b = 31; b >= 0; b--) {
            if (started) r = fp_sqr<P>(r);
            if ((e[i] >> b) & 1u) {
                r = started ? fp_mul<P>(r, a) : a;
                started = true;
            }
        }
    }
    return r;
}

template <class P>
CQB_HD Fp<P> fp_inv(const Fp<P>& a) {  // returns 0 for 0, like invert().unwrap_or(zero)
    uint32_t e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = P::mod(i);
    e[0] -= 2u;  // both moduli have low limb >= 2, no borrow
    return fp_pow<P>(a, e);
}

// Low-latency inversion for the places where ONE thread inverts ONE element on the critical path (the affine
// normalisation at the end of an MSM): binary extended Euclid on the 8-limb integers — ~20k simple ALU instructions instead
// of ~380 dependent Montgomery multiplications. Input/output in Montgomery form: for x = aR it returns a^-1 R
// (= x^-1 * R^2, obtained as mont_mul(x^-1, R^3)). Returns 0 for 0. Same value as fp_inv (the inverse is unique).
template <class P>
CQB_HD Fp<P> fp_inv_binary(const Fp<P>& a) {
    if (a.is_zero()) return a;
    uint32_t u[8], v[8], x1[8], x2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { u[i] = a.l[i]; v[i] = P::mod(i); x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    auto is_one = [](const uint32_t* w) {
        uint32_t o = w[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < 8; i++) o |= w[i];
        return o == 0;
    };
    auto shr1 = [](uint32_t* w) {
#pragma unroll
        for (int i = 0; i < 7; i++) w[i] = (w[i] >> 1) | (w[i + 1] << 31);
        w[7] >>= 1;
    };
    auto half_mod = [&](uint32_t* x) {  // x <- x/2 mod p  (x < p < 2^254, so x + p < 2^255 cannot overflow)
        if (x[0] & 1u) {
            uint64_t c = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) { c += (uint64_t)x[i] + P::mod(i); x[i] = (uint32_t)c; c >>= 32; }
        }
        shr1(x);
    };
    auto geq = [](const uint32_t* a_, const uint32_t* b_) {
        for (int i = 7; i >= 0; i--) {
            if (a_[i] != b_[i]) return a_[i] > b_[i];
        }
        return true;
    };
    auto sub = [](uint32_t* a_, const uint32_t* b_) {  // a -= b, returns borrow
        uint64_t br = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t d = (uint64_t)a_[i] - b_[i] - br;
            a_[i] = (uint32_t)d;
            br = (d >> 32) & 1u;
        }
        return (uint32_t)br;
    };
    auto sub_mod = [&](uint32_t* a_, const uint32_t* b_) {  // a <- a - b mod p
        if (sub(a_, b_)) {
            uint64_t c = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) { c += (uint64_t)a_[i] + P::mod(i); a_[i] = (uint32_t)c; c >>= 32; }
        }
    };
    while (!is_one(u) && !is_one(v)) {
        while (!(u[0] & 1u)) { shr1(u); half_mod(x1); }
        while (!(v[0] & 1u)) { shr1(v); half_mod(x2); }
        if (geq(u, v)) { sub(u, v); sub_mod(x1, x2); }
        else { sub(v, u); sub_mod(x2, x1); }
    }
    Fp<P> r;
    const bool from_u = is_one(u);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = from_u ? x1[i] : x2[i];
    return fp_mul<P>(r, Fp<P>::r3());
}

// Branch-free inversion for the places where EVERY lane of a warp inverts its own element at the same time (the batched-affine
// bucket accumulation: one inversion per thread per batch of additions). Fermat costs ~380 dependent modmuls on the
// multiplier pipe the additions themselves are bound by; the loops of fp_inv_binary diverge across lanes. This is the
// Bernstein-Yang "safegcd" iteration in its half-delta form over signed 30-bit limbs: 20 rounds of 30 division steps decided
// on the low limbs only (plain ALU work, uniform control flow), each round applied to f, g and to the Bezout coefficients
// d, e (mod p) as one 2x2 matrix (4 + 6 signed wide multiply-adds per limb). 600 steps suffice for any input below 2^256.
// Input/output in Montgomery form like fp_inv_binary: for x = aR it returns a^-1 R; 0 for 0. Same value as fp_inv.
template <class P>
struct SafeGcd {
    static constexpr uint32_t M30 = 0x3fffffffu;
    // limb j (30 bits) of the modulus
    static CQB_HD constexpr int32_t mod30(int j) {
        return (int32_t)((((uint64_t)P::mod((30 * j) / 32) | ((30 * j) / 32 + 1 < 8 ? (uint64_t)P::mod((30 * j) / 32 + 1) << 32 : 0ull)) >> ((30 * j) % 32)) & M30);
    }
    static CQB_HD constexpr uint32_t minv30() { return (0u - P::inv()) & M30; }  // p^-1 mod 2^30 (P::inv() is -p^-1 mod 2^32)
};

template <class P>
CQB_HD Fp<P> fp_inv_safegcd(const Fp<P>& a) {
    typedef SafeGcd<P> S;
    const int32_t M30 = (int32_t)S::M30;
    int32_t f[9], g[9], d[9], e[9];
#pragma unroll
    for (int j = 0; j < 9; j++) {
        f[j] = S::mod30(j);
        const int lo = (30 * j) / 32, sh = (30 * j) % 32;
        uint64_t w = a.l[lo];
        if (lo + 1 < 8) w |= (uint64_t)a.l[lo + 1] << 32;
        g[j] = (int32_t)((uint32_t)(w >> sh) & S::M30);
        d[j] = 0;
        e[j] = 0;
    }
    e[0] = 1;
    int32_t zeta = -1;  // -(delta + 1/2), delta = 1/2
#pragma unroll 1
    for (int round = 0; round < 20; round++) {
        // 30 division steps on the low limbs: the matrix [u v; q r] with [f'; g'] = 2^-30 [u v; q r] [f; g]
        uint32_t u = 1, v = 0, q = 0, r = 1, fl = (uint32_t)f[0], gl = (uint32_t)g[0];
#pragma unroll 6
        for (int i = 0; i < 30; i++) {
            uint32_t m1 = (uint32_t)(zeta >> 31);  // zeta < 0
            const uint32_t m2 = 0u - (gl & 1u);    // g odd
            const uint32_t x = (fl ^ m1) - m1, y = (u ^ m1) - m1, z = (v ^ m1) - m1;
            gl += x & m2;
            q += y & m2;
            r += z & m2;
            m1 &= m2;
            zeta = (int32_t)((uint32_t)zeta ^ m1) - 1;
            fl += gl & m1;
            u += q & m1;
            v += r & m1;
            gl >>= 1;
            u <<= 1;
            v <<= 1;
        }
        const int32_t su = (int32_t)u, sv = (int32_t)v, sq = (int32_t)q, sr = (int32_t)r;
        {   // [d; e] <- 2^-30 [u v; q r] [d; e] mod p, kept in (-2p, p)
            const int32_t sd = d[8] >> 31, se = e[8] >> 31;
            int32_t md = (su & sd) + (sv & se), me = (sq & sd) + (sr & se);
            int64_t cd = (int64_t)su * d[0] + (int64_t)sv * e[0];
            int64_t ce = (int64_t)sq * d[0] + (int64_t)sr * e[0];
            md -= (int32_t)((S::minv30() * (uint32_t)cd + (uint32_t)md) & S::M30);
            me -= (int32_t)((S::minv30() * (uint32_t)ce + (uint32_t)me) & S::M30);
            cd += (int64_t)S::mod30(0) * md;
            ce += (int64_t)S::mod30(0) * me;
            cd >>= 30;
            ce >>= 30;
#pragma unroll
            for (int i = 1; i < 9; i++) {
                const int32_t di = d[i], ei = e[i];
                cd += (int64_t)su * di + (int64_t)sv * ei;
                ce += (int64_t)sq * di + (int64_t)sr * ei;
                cd += (int64_t)S::mod30(i) * md;
                ce += (int64_t)S::mod30(i) * me;
                d[i - 1] = (int32_t)cd & M30;
                cd >>= 30;
                e[i - 1] = (int32_t)ce & M30;
                ce >>= 30;
            }
            d[8] = (int32_t)cd;
            e[8] = (int32_t)ce;
        }
        {   // [f; g] <- 2^-30 [u v; q r] [f; g] (exact: the low 30 bits vanish by construction)
            int64_t cf = (int64_t)su * f[0] + (int64_t)sv * g[0];
            int64_t cg = (int64_t)sq * f[0] + (int64_t)sr * g[0];
            cf >>= 30;
            cg >>= 30;
#pragma unroll
            for (int i = 1; i < 9; i++) {
                const int32_t fi = f[i], gi = g[i];
                cf += (int64_t)su * fi + (int64_t)sv * gi;
                cg += (int64_t)sq * fi + (int64_t)sr * gi;
                f[i - 1] = (int32_t)cf & M30;
                cf >>= 30;
                g[i - 1] = (int32_t)cg & M30;
                cg >>= 30;
            }
            f[8] = (int32_t)cf;
            g[8] = (int32_t)cg;
        }
    }
    // now g = 0 and f = +-1 (for a != 0): the inverse is sign(f) * d, brought from (-2p, p) into [0, p)
    {
        int32_t add = d[8] >> 31;
#pragma unroll
        for (int j = 0; j < 9; j++) d[j] += S::mod30(j) & add;
        const int32_t neg = f[8] >> 31;
#pragma unroll
        for (int j = 0; j < 9; j++) d[j] = (d[j] ^ neg) - neg;
#pragma unroll
        for (int j = 0; j < 8; j++) { d[j + 1] += d[j] >> 30; d[j] &= M30; }
        add = d[8] >> 31;
#pragma unroll
        for (int j = 0; j < 9; j++) d[j] += S::mod30(j) & add;
#pragma unroll
        for (int j = 0; j < 8; j++) { d[j + 1] += d[j] >> 30; d[j] &= M30; }
    }
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < 8; i++) {  // 9 x 30-bit limbs -> 8 x 32-bit limbs
        const int lo = (32 * i) / 30, sh = (32 * i) % 30;
        uint64_t w = (uint64_t)(uint32_t)d[lo] | ((uint64_t)(uint32_t)d[lo + 1] << 30);
        if (lo + 2 < 9) w |= (uint64_t)(uint32_t)d[lo + 2] << 60;
        r.l[i] = (uint32_t)(w >> sh);
    }
    return fp_mul<P>(r, Fp<P>::r3());
}

typedef Fp<FrP> Fr;
typedef Fp<FqP> Fq;

}  // namespace cqb

// ec.cuh — bn256 G1 (y^2 = x^3 + 3 over Fq) point arithmetic for the MSM kernels.
//
// Parity contract: the reference computes in Jacobian coordinates (arithmetic/curves/src/derive/curve.rs:422-447 double,
// :809-851 add, :853-893 mixed add) and only the AFFINE normal form is canonical (to_affine :399-412).
// This file therefore uses the coordinate system that is cheapest on the B200 integer pipe — extended Jacobian "XYZZ"
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), 8M+2S for a mixed add instead of the reference's 7M+4S — and handles the same
// exceptional cases the reference branches on (identity operands, P+P, P+(-P): curve.rs:818-824, 866-871).
// Identity conventions match the reference: affine identity = (0,0) (curve.rs:696-709), projective identity has ZZ = 0.
#pragma once
#include "fp.cuh"

namespace cqb {

struct G1Affine {
    Fq x, y;
    CQB_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

struct G1Xyzz {
    Fq x, y, zz, zzz;
    static CQB_HD G1Xyzz identity() {
        G1Xyzz p;
        p.x = Fq::zero(); p.y = Fq::zero(); p.zz = Fq::zero(); p.zzz = Fq::zero();
        return p;
    }
    CQB_HD bool is_identity() const { return zz.is_zero(); }
    static CQB_HD G1Xyzz from_affine(const G1Affine& a) {
        G1Xyzz p;
        if (a.is_identity()) return identity();
        p.x = a.x; p.y = a.y; p.zz = Fq::one(); p.zzz = Fq::one();
        return p;
    }
};

// Multiplier policies. FqInline expands every multiplication in place (straight-line code, ~2.8 KB per modmul): right for
// the throughput kernels, where several warps per scheduler walk one hot loop. FqCall routes them through ONE non-inlined
// copy per kernel: the latency-bound tail kernels (bucket merge / reduction / final sums: chains of dependent XYZZ
// operations executed by one warp per scheduler) otherwise run out of the instruction caches — a single inlined g1_add
// is ~40 KB of code — and stall on instruction fetch.
struct FqInline {
    static CQB_HD Fq mul(const Fq& a, const Fq& b) { return fp_mul<FqP>(a, b); }
    static CQB_HD Fq sqr(const Fq& a) { return fp_sqr<FqP>(a); }
    static CQB_HD Fq msub(const Fq& a, const Fq& b, const Fq& c, const Fq& d) { return fp_mul2<FqP>(a, b, fp_neg<FqP>(c), d); }
};
#if defined(__CUDACC__)
struct FqCall {
    static __device__ __noinline__ Fq mul(Fq a, Fq b) { return fp_mul<FqP>(a, b); }
    static __device__ __forceinline__ Fq sqr(const Fq& a) { return mul(a, a); }
    static __device__ __noinline__ Fq msub(Fq a, Fq b, Fq c, Fq d) { return fp_mul2<FqP>(a, b, fp_neg<FqP>(c), d); }
};
#else
typedef FqInline FqCall;
#endif

#define FQM(a, b) M::mul(a, b)
#define FQS(a) M::sqr(a)
#define FQA(a, b) fp_add<FqP>(a, b)
#define FQSUB(a, b) fp_sub<FqP>(a, b)
// a*b - c*d with one Montgomery reduction (the negation is 8 ALU ops; the fused product saves 72 wide multiplies)
#define FQMSUB(a, b, c, d) M::msub(a, b, c, d)

// 2*(x,y) for an affine, non-identity point (mdbl-2008-s-1). y == 0 cannot happen on this curve (no 2-torsion), but
// the formula degrades gracefully to ZZ = 0 = identity anyway.
template <class M = FqInline>
CQB_HD G1Xyzz g1_double_affine(const Fq& x1, const Fq& y1) {
    G1Xyzz r;
    Fq u = fp_dbl<FqP>(y1);
    Fq v = FQS(u);
    Fq w = FQM(u, v);
    Fq s = FQM(x1, v);
    Fq xx = FQS(x1);
    Fq m = FQA(fp_dbl<FqP>(xx), xx);
    r.x = FQSUB(FQS(m), fp_dbl<FqP>(s));
    r.y = FQMSUB(m, FQSUB(s, r.x), y1, w);
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2*P (dbl-2008-s-1)
template <class M = FqInline>
CQB_HD G1Xyzz g1_double(const G1Xyzz& p) {
    if (p.is_identity()) return p;
    G1Xyzz r;
    Fq u = fp_dbl<FqP>(p.y);
    Fq v = FQS(u);
    Fq w = FQM(u, v);
    Fq s = FQM(p.x, v);
    Fq xx = FQS(p.x);
    Fq m = FQA(fp_dbl<FqP>(xx), xx);
    r.x = FQSUB(FQS(m), fp_dbl<FqP>(s));
    r.y = FQMSUB(m, FQSUB(s, r.x), p.y, w);
    r.zz = FQM(v, p.zz);
    r.zzz = FQM(w, p.zzz);
    return r;
}

// acc += (x2, y2) with (x2,y2) affine and NOT the identity (madd-2008-s); all exceptional cases handled.
template <class M = FqInline>
CQB_HD void g1_madd(G1Xyzz& acc, const Fq& x2, const Fq& y2) {
    if (acc.is_identity()) {
        acc.x = x2; acc.y = y2; acc.zz = Fq::one(); acc.zzz = Fq::one();
        return;
    }
    Fq u2 = FQM(x2, acc.zz);
    Fq s2 = FQM(y2, acc.zzz);
    Fq p = FQSUB(u2, acc.x);
    Fq r = FQSUB(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) acc = g1_double_affine<M>(x2, y2);  // same point: reference curve.rs:866-868
        else acc = G1Xyzz::identity();                    // opposite points: curve.rs:869-870
        return;
    }
    Fq pp = FQS(p);
    Fq ppp = FQM(p, pp);
    Fq q = FQM(acc.x, pp);
    Fq x3 = FQSUB(FQSUB(FQS(r), ppp), fp_dbl<FqP>(q));
    Fq y3 = FQMSUB(r, FQSUB(q, x3), acc.y, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = FQM(acc.zz, pp);
    acc.zzz = FQM(acc.zzz, ppp);
}

// acc += b (add-2008-s), all exceptional cases handled.
template <class M = FqInline>
CQB_HD void g1_add(G1Xyzz& acc, const G1Xyzz& b) {
    if (b.is_identity()) return;
    if (acc.is_identity()) { acc = b; return; }
    Fq u1 = FQM(acc.x, b.zz);
    Fq u2 = FQM(b.x, acc.zz);
    Fq s1 = FQM(acc.y, b.zzz);
    Fq s2 = FQM(b.y, acc.zzz);
    Fq p = FQSUB(u2, u1);
    Fq r = FQSUB(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) acc = g1_double<M>(acc);  // reference curve.rs:818-820
        else acc = G1Xyzz::identity();          // curve.rs:821-823
        return;
    }
    Fq pp = FQS(p);
    Fq ppp = FQM(p, pp);
    Fq q = FQM(u1, pp);
    Fq x3 = FQSUB(FQSUB(FQS(r), ppp), fp_dbl<FqP>(q));
    Fq y3 = FQMSUB(r, FQSUB(q, x3), s1, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = FQM(FQM(acc.zz, b.zz), pp);
    acc.zzz = FQM(FQM(acc.zzz, b.zzz), ppp);
}

// XYZZ -> affine normal form (what the reference's to_affine / batch_normalize produce, curve.rs:399-412): one inversion.
template <class M = FqInline>
CQB_HD G1Affine g1_to_affine(const G1Xyzz& p) {
    G1Affine a;
    if (p.is_identity()) { a.x = Fq::zero(); a.y = Fq::zero(); return a; }
    Fq inv = fp_inv<FqP>(FQM(p.zz, p.zzz));
    Fq zz_inv = FQM(inv, p.zzz);
    Fq zzz_inv = FQM(inv, p.zz);
    a.x = FQM(p.x, zz_inv);
    a.y = FQM(p.y, zzz_inv);
    return a;
}

// same result through the low-latency binary inversion: for kernels where a single thread normalises a single point
template <class M = FqInline>
CQB_HD G1Affine g1_to_affine_lowlat(const G1Xyzz& p) {
    G1Affine a;
    if (p.is_identity()) { a.x = Fq::zero(); a.y = Fq::zero(); return a; }
    Fq inv = fp_inv_safegcd<FqP>(FQM(p.zz, p.zzz));  // 22 us single-thread latency (42.9 k cycles) against ~40 us for the binary Euclid loop
    Fq zz_inv = FQM(inv, p.zzz);
    Fq zzz_inv = FQM(inv, p.zz);
    a.x = FQM(p.x, zz_inv);
    a.y = FQM(p.y, zzz_inv);
    return a;
}

#undef FQM
#undef FQS
#undef FQA
#undef FQSUB
#undef FQMSUB

}  // namespace cqb

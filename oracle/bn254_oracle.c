/* bn254_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle for the parity tests, smoke() and bench.py's cpu_baseline /
 * --impl reference legs). Nothing in the product path (sha2-on-cq-halo2_b200/) links, imports or calls this file.
 *
 * It is a plain-C restatement of the reference's OWN algorithms for the hot path, function by function, each citing
 * the reference file:line it follows (reference = aleph-zero-foundation/sha2-on-cq-halo2, Rust, which cannot be
 * compiled in this image: no cargo/rustc, no Cargo.lock, no vendored registry — SURVEY.md F8):
 *
 *   Fr / Fq          arithmetic/curves/src/derive/field.rs + bn256/{fr,fq}.rs        (field_impl.inc)
 *   G1               arithmetic/curves/src/derive/curve.rs, bn256/curve.rs
 *   MSM / FFT        halo2_proofs/src/arithmetic.rs:13-159, 171-274
 *   domain wrappers  halo2_proofs/src/poly/domain.rs:39-142, 238-374
 *   KZG params       halo2_proofs/src/poly/kzg/commitment.rs:71-178, 209-276, 496-504, 539-543
 *   CQ sparse sums   halo2_proofs/src/plonk/static_lookup/prover.rs:167-170, 245-257
 *
 * PARITY PINNING: the reference holds NO golden vectors for MSM / FFT / commitments / proofs. What
 * it does hold for this path — field constants and KATs (fr.rs:320-367, fq.rs:331-351), curve laws (tests/curve.rs),
 * commit(ifft(a)) == commit_lagrange(a) (kzg/commitment.rs:570-593) — is checked against this oracle by
 * tests/test_oracle.py, together with an independent Python big-integer model (oracle/pyref.py). MSM/FFT outputs are
 * therefore "parity unpinned by fixtures": pinned by restatement + mathematical uniqueness (one canonical affine point,
 * one canonical limb vector) only.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;
#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------------------------
 * Fr — bn256/fr.rs:29-118
 * ------------------------------------------------------------------------------------------------------------------ */
#define FN(x) fr_##x
#define F_MOD0 0x43e1f593f0000001ULL
#define F_MOD1 0x2833e84879b97091ULL
#define F_MOD2 0xb85045b68181585dULL
#define F_MOD3 0x30644e72e131a029ULL
#define F_INV 0xc2e1f593efffffffULL
static const fe FR_R = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const fe FR_R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
static const fe FR_R3 = {{0x5e94d8e1b4bf0040ULL, 0x2a489cbe1cfbb6b8ULL, 0x893cc664a19fcfedULL, 0x0cf8594b7fcc657cULL}};
#define F_R FR_R
#define F_R2 FR_R2
#define F_R3 FR_R3
#include "field_impl.inc"
#undef FN
#undef F_MOD0
#undef F_MOD1
#undef F_MOD2
#undef F_MOD3
#undef F_INV
#undef F_R
#undef F_R2
#undef F_R3

/* ------------------------------------------------------------------------------------------------------------------
 * Fq — bn256/fq.rs:28-90
 * ------------------------------------------------------------------------------------------------------------------ */
#define FN(x) fq_##x
#define F_MOD0 0x3c208c16d87cfd47ULL
#define F_MOD1 0x97816a916871ca8dULL
#define F_MOD2 0xb85045b68181585dULL
#define F_MOD3 0x30644e72e131a029ULL
#define F_INV 0x87d20782e4866389ULL
static const fe FQ_R = {{0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL}};
static const fe FQ_R2 = {{0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL}};
static const fe FQ_R3 = {{0xb1cd6dafda1530dfULL, 0x62f210e6a7283db6ULL, 0xef7f0b0c0ada0afbULL, 0x20fd6e902d592544ULL}};
#define F_R FQ_R
#define F_R2 FQ_R2
#define F_R3 FQ_R3
#include "field_impl.inc"
#undef FN

/* Fr constants that are from_raw(...) in the reference (fr.rs:70-118) */
static fe fr_root_of_unity(void) { return fr_from_raw(0xd34f1ed960c37c9cULL, 0x3215cf6dd39329c8ULL, 0x98865ea93dd31f74ULL, 0x03ddb9f5166d18b7ULL); }
static fe fr_root_of_unity_inv(void) { return fr_from_raw(0x0ed3e50a414e6dbaULL, 0xb22625f59115aba7ULL, 0x1bbe587180f34361ULL, 0x048127174daabc26ULL); }
static fe fr_two_inv(void) { return fr_from_raw(0xa1f0fac9f8000001ULL, 0x9419f4243cdcb848ULL, 0xdc2822db40c0ac2eULL, 0x183227397098d014ULL); }
static fe fr_delta(void) { return fr_from_raw(0x870e56bbe533e9a2ULL, 0x5b5f898e5e963f25ULL, 0x64ec26aad4c86e71ULL, 0x09226b6e22c6f0caULL); }
static fe fr_zeta(void) { return fr_from_raw(0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL); }
#define FR_S 28 /* fr.rs:72 */

/* ------------------------------------------------------------------------------------------------------------------
 * G1 — derive/curve.rs (new_curve_impl!), instantiated bn256/curve.rs:23-34 with generator (1,2), b = 3 (:66-68)
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct { fe x, y; } g1a;    /* affine; identity = (0,0)  curve.rs:696-705 */
typedef struct { fe x, y, z; } g1j; /* Jacobian; identity z = 0  curve.rs:453-463 */

static g1j g1j_identity(void) { g1j p; p.x = fq_zero(); p.y = fq_zero(); p.z = fq_zero(); return p; }
static g1a g1a_identity(void) { g1a p; p.x = fq_zero(); p.y = fq_zero(); return p; }
static int g1j_is_identity(const g1j* p) { return fq_is_zero(p->z); }
static int g1a_is_identity(const g1a* p) { return fq_is_zero(p->x) & fq_is_zero(p->y); } /* curve.rs:707-709 */
static g1a g1a_generator(void) { g1a g; g.x = fq_one(); g.y = fq_from_raw(2, 0, 0, 0); return g; }
static fe g1_b(void) { return fq_from_raw(3, 0, 0, 0); }

/* curve.rs:711-717 to_curve */
static g1j g1a_to_curve(const g1a* p) {
    g1j r; r.x = p->x; r.y = p->y; r.z = g1a_is_identity(p) ? fq_zero() : fq_one(); return r;
}
/* curve.rs:1044ff Neg for affine: (x, -y) */
static g1a g1a_neg(const g1a* p) { g1a r; r.x = p->x; r.y = fq_neg(p->y); return r; }
static g1j g1j_neg(const g1j* p) { g1j r = *p; r.y = fq_neg(p->y); return r; }

/* curve.rs:422-447 double */
static g1j g1j_double(const g1j* p) {
    fe a = fq_square(p->x);
    fe b = fq_square(p->y);
    fe c = fq_square(b);
    fe d = fq_add(p->x, b);
    d = fq_square(d);
    d = fq_sub(fq_sub(d, a), c);
    d = fq_add(d, d);
    fe e = fq_add(fq_add(a, a), a);
    fe f = fq_square(e);
    fe z3 = fq_mul(p->z, p->y);
    z3 = fq_add(z3, z3);
    fe x3 = fq_sub(f, fq_add(d, d));
    c = fq_add(c, c);
    c = fq_add(c, c);
    c = fq_add(c, c);
    fe y3 = fq_sub(fq_mul(e, fq_sub(d, x3)), c);
    g1j r; r.x = x3; r.y = y3; r.z = z3;
    if (g1j_is_identity(p)) return g1j_identity();
    return r;
}

/* curve.rs:809-851 Jacobian + Jacobian */
static g1j g1j_add(const g1j* s, const g1j* rhs) {
    if (g1j_is_identity(s)) return *rhs;
    if (g1j_is_identity(rhs)) return *s;
    fe z1z1 = fq_square(s->z);
    fe z2z2 = fq_square(rhs->z);
    fe u1 = fq_mul(s->x, z2z2);
    fe u2 = fq_mul(rhs->x, z1z1);
    fe s1 = fq_mul(fq_mul(s->y, z2z2), rhs->z);
    fe s2 = fq_mul(fq_mul(rhs->y, z1z1), s->z);
    if (fq_eq(u1, u2)) {
        if (fq_eq(s1, s2)) return g1j_double(s);
        return g1j_identity();
    }
    fe h = fq_sub(u2, u1);
    fe i = fq_square(fq_add(h, h));
    fe j = fq_mul(h, i);
    fe r = fq_sub(s2, s1);
    r = fq_add(r, r);
    fe v = fq_mul(u1, i);
    fe x3 = fq_sub(fq_sub(fq_sub(fq_square(r), j), v), v);
    s1 = fq_mul(s1, j);
    s1 = fq_add(s1, s1);
    fe y3 = fq_sub(fq_mul(r, fq_sub(v, x3)), s1);
    fe z3 = fq_sub(fq_sub(fq_square(fq_add(s->z, rhs->z)), z1z1), z2z2);
    z3 = fq_mul(z3, h);
    g1j o; o.x = x3; o.y = y3; o.z = z3;
    return o;
}

/* curve.rs:853-893 Jacobian + affine (mixed) */
static g1j g1j_madd(const g1j* s, const g1a* rhs) {
    if (g1j_is_identity(s)) return g1a_to_curve(rhs);
    if (g1a_is_identity(rhs)) return *s;
    fe z1z1 = fq_square(s->z);
    fe u2 = fq_mul(rhs->x, z1z1);
    fe s2 = fq_mul(fq_mul(rhs->y, z1z1), s->z);
    if (fq_eq(s->x, u2)) {
        if (fq_eq(s->y, s2)) return g1j_double(s);
        return g1j_identity();
    }
    fe h = fq_sub(u2, s->x);
    fe hh = fq_square(h);
    fe i = fq_add(hh, hh);
    i = fq_add(i, i);
    fe j = fq_mul(h, i);
    fe r = fq_sub(s2, s->y);
    r = fq_add(r, r);
    fe v = fq_mul(s->x, i);
    fe x3 = fq_sub(fq_sub(fq_sub(fq_square(r), j), v), v);
    j = fq_mul(s->y, j);
    j = fq_add(j, j);
    fe y3 = fq_sub(fq_mul(r, fq_sub(v, x3)), j);
    fe z3 = fq_sub(fq_sub(fq_square(fq_add(s->z, h)), z1z1), hh);
    g1j o; o.x = x3; o.y = y3; o.z = z3;
    return o;
}

/* curve.rs:964-1000 affine + affine -> Jacobian */
static g1j g1a_add(const g1a* s, const g1a* rhs) {
    if (g1a_is_identity(s)) return g1a_to_curve(rhs);
    if (g1a_is_identity(rhs)) return g1a_to_curve(s);
    if (fq_eq(s->x, rhs->x)) {
        if (fq_eq(s->y, rhs->y)) { g1j t = g1a_to_curve(s); return g1j_double(&t); }
        return g1j_identity();
    }
    fe h = fq_sub(rhs->x, s->x);
    fe hh = fq_square(h);
    fe i = fq_add(hh, hh);
    i = fq_add(i, i);
    fe j = fq_mul(h, i);
    fe r = fq_sub(rhs->y, s->y);
    r = fq_add(r, r);
    fe v = fq_mul(s->x, i);
    fe x3 = fq_sub(fq_sub(fq_sub(fq_square(r), j), v), v);
    j = fq_mul(s->y, j);
    j = fq_add(j, j);
    fe y3 = fq_sub(fq_mul(r, fq_sub(v, x3)), j);
    fe z3 = fq_add(h, h);
    g1j o; o.x = x3; o.y = y3; o.z = z3;
    return o;
}

/* curve.rs:399-412 to_affine */
static g1a g1j_to_affine(const g1j* p) {
    fe zinv = fq_invert(p->z); /* invert().unwrap_or(zero): pow maps 0 -> 0 */
    fe zinv2 = fq_square(zinv);
    fe x = fq_mul(p->x, zinv2);
    fe zinv3 = fq_mul(zinv2, zinv);
    fe y = fq_mul(p->y, zinv3);
    if (fq_is_zero(zinv)) return g1a_identity();
    g1a r; r.x = x; r.y = y;
    return r;
}

/* curve.rs:362-397 batch_normalize (Montgomery's trick; identities skipped) */
static void g1j_batch_normalize(const g1j* p, g1a* q, size_t n) {
    fe acc = fq_one();
    for (size_t i = 0; i < n; i++) {
        q[i].x = acc;
        if (!g1j_is_identity(&p[i])) acc = fq_mul(acc, p[i].z);
    }
    acc = fq_invert(acc);
    for (size_t k = n; k-- > 0;) {
        int skip = g1j_is_identity(&p[k]);
        fe tmp = fq_mul(q[k].x, acc);
        if (!skip) acc = fq_mul(acc, p[k].z);
        fe tmp2 = fq_square(tmp);
        fe tmp3 = fq_mul(tmp2, tmp);
        q[k].x = fq_mul(p[k].x, tmp2);
        q[k].y = fq_mul(p[k].y, tmp3);
        if (skip) q[k] = g1a_identity();
    }
}

/* curve.rs:914-935 / 1019-1040: double-and-add over to_repr() bits, MSB first (used by `G1Affine * Fr`, `G1 * Fr`) */
static g1j g1j_mul(const g1j* p, fe scalar) {
    uint8_t repr[32];
    fr_to_repr(scalar, repr);
    g1j acc = g1j_identity();
    for (int byte = 31; byte >= 0; byte--)
        for (int i = 7; i >= 0; i--) {
            acc = g1j_double(&acc);
            if ((repr[byte] >> i) & 1) acc = g1j_add(&acc, p);
        }
    return acc;
}
static g1j g1a_mul(const g1a* p, fe scalar) {
    uint8_t repr[32];
    fr_to_repr(scalar, repr);
    g1j acc = g1j_identity();
    for (int byte = 31; byte >= 0; byte--)
        for (int i = 7; i >= 0; i--) {
            acc = g1j_double(&acc);
            if ((repr[byte] >> i) & 1) acc = g1j_madd(&acc, p);
        }
    return acc;
}

/* curve.rs:635-646 compressed to_bytes: canonical x, y-parity in bit 255; identity = 32 zero bytes */
static void g1a_to_bytes(const g1a* p, uint8_t out[32]) {
    if (g1a_is_identity(p)) { memset(out, 0, 32); return; }
    uint8_t yb[32];
    fq_to_repr(p->y, yb);
    fq_to_repr(p->x, out);
    out[31] |= (uint8_t)((yb[0] & 1) << 7);
}

/* curve.rs CurveExt::is_on_curve for affine: y^2 = x^3 + b, or identity */
static int g1a_is_on_curve(const g1a* p) {
    if (g1a_is_identity(p)) return 1;
    fe lhs = fq_square(p->y);
    fe rhs = fq_add(fq_mul(fq_square(p->x), p->x), g1_b());
    return fq_eq(lhs, rhs);
}

/* ------------------------------------------------------------------------------------------------------------------
 * Fq2 = Fq[u]/(u^2 + 1) — arithmetic/curves/src/bn256/fq2.rs:161-307 — and G2 — derive/curve.rs (new_curve_impl!) instantiated
 * bn256/curve.rs:36-48 over Fq2 with G2_GENERATOR_X / _Y (:100-129) and b = G2_B = 3/(9+u) (:85-98). The same generic formulas
 * as G1 above (the macro is generic in the base field), restated over fq2.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct { fe c0, c1; } fe2;
static fe2 fq2_zero(void) { fe2 r; r.c0 = fq_zero(); r.c1 = fq_zero(); return r; }
static fe2 fq2_one(void) { fe2 r; r.c0 = fq_one(); r.c1 = fq_zero(); return r; }
static int fq2_is_zero(fe2 a) { return fq_is_zero(a.c0) & fq_is_zero(a.c1); }
static int fq2_eq(fe2 a, fe2 b) { return fq_eq(a.c0, b.c0) & fq_eq(a.c1, b.c1); }
static fe2 fq2_add(fe2 a, fe2 b) { fe2 r; r.c0 = fq_add(a.c0, b.c0); r.c1 = fq_add(a.c1, b.c1); return r; }  /* fq2.rs:195-200 */
static fe2 fq2_sub(fe2 a, fe2 b) { fe2 r; r.c0 = fq_sub(a.c0, b.c0); r.c1 = fq_sub(a.c1, b.c1); return r; }  /* :202-207 */
static fe2 fq2_neg(fe2 a) { fe2 r; r.c0 = fq_neg(a.c0); r.c1 = fq_neg(a.c1); return r; }                      /* :221-226 */
static fe2 fq2_mul(fe2 a, fe2 b) { /* :161-170 */
    fe t1 = fq_mul(a.c0, b.c0);
    fe t0 = fq_add(a.c0, a.c1);
    fe t2 = fq_mul(a.c1, b.c1);
    fe s = fq_add(b.c0, b.c1);
    fe2 r;
    r.c0 = fq_sub(t1, t2);
    t1 = fq_add(t1, t2);
    t0 = fq_mul(t0, s);
    r.c1 = fq_sub(t0, t1);
    return r;
}
static fe2 fq2_square(fe2 a) { /* :172-181 */
    fe ab = fq_mul(a.c0, a.c1);
    fe c0c1 = fq_add(a.c0, a.c1);
    fe c0 = fq_add(fq_neg(a.c1), a.c0);
    c0 = fq_mul(c0, c0c1);
    c0 = fq_sub(c0, ab);
    fe2 r;
    r.c1 = fq_add(ab, ab);
    r.c0 = fq_add(c0, ab);
    return r;
}
static fe2 fq2_invert(fe2 a) { /* :290-307; 0 -> 0 (unwrap_or(zero) at the call sites) */
    fe t = fq_add(fq_square(a.c0), fq_square(a.c1));
    t = fq_invert(t);
    fe2 r;
    r.c0 = fq_mul(a.c0, t);
    r.c1 = fq_neg(fq_mul(a.c1, t));
    return r;
}
typedef struct { fe2 x, y; } g2a;    /* affine, 128 bytes: x.c0 x.c1 y.c0 y.c1; identity = zeros */
typedef struct { fe2 x, y, z; } g2j; /* Jacobian; identity z = 0 */
static g2j g2j_identity(void) { g2j p; p.x = fq2_zero(); p.y = fq2_zero(); p.z = fq2_zero(); return p; }
static g2a g2a_identity(void) { g2a p; p.x = fq2_zero(); p.y = fq2_zero(); return p; }
static int g2j_is_identity(const g2j* p) { return fq2_is_zero(p->z); }
static int g2a_is_identity(const g2a* p) { return fq2_is_zero(p->x) & fq2_is_zero(p->y); }
static g2a g2a_generator(void) { /* bn256/curve.rs:100-129 */
    g2a g;
    g.x.c0 = fq_from_raw(0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL);
    g.x.c1 = fq_from_raw(0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL);
    g.y.c0 = fq_from_raw(0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL);
    g.y.c1 = fq_from_raw(0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL);
    return g;
}
static fe2 g2_b(void) { /* bn256/curve.rs:85-98 */
    fe2 b;
    b.c0 = fq_from_raw(0x3267e6dc24a138e5ULL, 0xb5b4c5e559dbefa3ULL, 0x81be18991be06ac3ULL, 0x2b149d40ceb8aaaeULL);
    b.c1 = fq_from_raw(0xe4a2bd0685c315d2ULL, 0xa74fa084e52d1852ULL, 0xcd2cafadeed8fdf4ULL, 0x009713b03af0fed4ULL);
    return b;
}
static g2j g2a_to_curve(const g2a* p) { g2j r; r.x = p->x; r.y = p->y; r.z = g2a_is_identity(p) ? fq2_zero() : fq2_one(); return r; }
static g2a g2a_neg(const g2a* p) { g2a r; r.x = p->x; r.y = fq2_neg(p->y); return r; }
static g2j g2j_double(const g2j* p) { /* derive/curve.rs:422-447 */
    fe2 a = fq2_square(p->x);
    fe2 b = fq2_square(p->y);
    fe2 c = fq2_square(b);
    fe2 d = fq2_square(fq2_add(p->x, b));
    d = fq2_sub(fq2_sub(d, a), c);
    d = fq2_add(d, d);
    fe2 e = fq2_add(fq2_add(a, a), a);
    fe2 f = fq2_square(e);
    fe2 z3 = fq2_mul(p->z, p->y);
    z3 = fq2_add(z3, z3);
    fe2 x3 = fq2_sub(f, fq2_add(d, d));
    c = fq2_add(c, c);
    c = fq2_add(c, c);
    c = fq2_add(c, c);
    fe2 y3 = fq2_sub(fq2_mul(e, fq2_sub(d, x3)), c);
    g2j r; r.x = x3; r.y = y3; r.z = z3;
    if (g2j_is_identity(p)) return g2j_identity();
    return r;
}
static g2j g2j_add(const g2j* s, const g2j* rhs) { /* :809-851 */
    if (g2j_is_identity(s)) return *rhs;
    if (g2j_is_identity(rhs)) return *s;
    fe2 z1z1 = fq2_square(s->z);
    fe2 z2z2 = fq2_square(rhs->z);
    fe2 u1 = fq2_mul(s->x, z2z2);
    fe2 u2 = fq2_mul(rhs->x, z1z1);
    fe2 s1 = fq2_mul(fq2_mul(s->y, z2z2), rhs->z);
    fe2 s2 = fq2_mul(fq2_mul(rhs->y, z1z1), s->z);
    if (fq2_eq(u1, u2)) {
        if (fq2_eq(s1, s2)) return g2j_double(s);
        return g2j_identity();
    }
    fe2 h = fq2_sub(u2, u1);
    fe2 i = fq2_square(fq2_add(h, h));
    fe2 j = fq2_mul(h, i);
    fe2 r = fq2_sub(s2, s1);
    r = fq2_add(r, r);
    fe2 v = fq2_mul(u1, i);
    fe2 x3 = fq2_sub(fq2_sub(fq2_sub(fq2_square(r), j), v), v);
    s1 = fq2_mul(s1, j);
    s1 = fq2_add(s1, s1);
    fe2 y3 = fq2_sub(fq2_mul(r, fq2_sub(v, x3)), s1);
    fe2 z3 = fq2_sub(fq2_sub(fq2_square(fq2_add(s->z, rhs->z)), z1z1), z2z2);
    z3 = fq2_mul(z3, h);
    g2j o; o.x = x3; o.y = y3; o.z = z3;
    return o;
}
static g2j g2j_madd(const g2j* s, const g2a* rhs) { /* :853-893 */
    if (g2j_is_identity(s)) return g2a_to_curve(rhs);
    if (g2a_is_identity(rhs)) return *s;
    fe2 z1z1 = fq2_square(s->z);
    fe2 u2 = fq2_mul(rhs->x, z1z1);
    fe2 s2 = fq2_mul(fq2_mul(rhs->y, z1z1), s->z);
    if (fq2_eq(s->x, u2)) {
        if (fq2_eq(s->y, s2)) return g2j_double(s);
        return g2j_identity();
    }
    fe2 h = fq2_sub(u2, s->x);
    fe2 hh = fq2_square(h);
    fe2 i = fq2_add(hh, hh);
    i = fq2_add(i, i);
    fe2 j = fq2_mul(h, i);
    fe2 r = fq2_sub(s2, s->y);
    r = fq2_add(r, r);
    fe2 v = fq2_mul(s->x, i);
    fe2 x3 = fq2_sub(fq2_sub(fq2_sub(fq2_square(r), j), v), v);
    j = fq2_mul(s->y, j);
    j = fq2_add(j, j);
    fe2 y3 = fq2_sub(fq2_mul(r, fq2_sub(v, x3)), j);
    fe2 z3 = fq2_sub(fq2_sub(fq2_square(fq2_add(s->z, h)), z1z1), hh);
    g2j o; o.x = x3; o.y = y3; o.z = z3;
    return o;
}
static g2a g2j_to_affine(const g2j* p) { /* :399-412 */
    fe2 zinv = fq2_invert(p->z);
    fe2 zinv2 = fq2_square(zinv);
    fe2 x = fq2_mul(p->x, zinv2);
    fe2 y = fq2_mul(p->y, fq2_mul(zinv2, zinv));
    if (fq2_is_zero(zinv)) return g2a_identity();
    g2a r; r.x = x; r.y = y;
    return r;
}
static g2j g2a_mul(const g2a* p, fe scalar) { /* :1019-1040 */
    uint8_t repr[32];
    fr_to_repr(scalar, repr);
    g2j acc = g2j_identity();
    for (int byte = 31; byte >= 0; byte--)
        for (int i = 7; i >= 0; i--) {
            acc = g2j_double(&acc);
            if ((repr[byte] >> i) & 1) acc = g2j_madd(&acc, p);
        }
    return acc;
}
static int g2a_is_on_curve(const g2a* p) {
    if (g2a_is_identity(p)) return 1;
    return fq2_eq(fq2_square(p->y), fq2_add(fq2_mul(fq2_square(p->x), p->x), g2_b()));
}

/* ------------------------------------------------------------------------------------------------------------------
 * MSM — halo2_proofs/src/arithmetic.rs:13-159
 * ------------------------------------------------------------------------------------------------------------------ */
/* arithmetic.rs:24-42 get_at */
static size_t msm_get_at(size_t segment, size_t c, const uint8_t bytes[32]) {
    size_t skip_bits = segment * c;
    size_t skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0};
    for (size_t i = 0; i < 8 && skip_bytes + i < 32; i++) v[i] = bytes[skip_bytes + i];
    uint64_t tmp = 0;
    for (int i = 7; i >= 0; i--) tmp = (tmp << 8) | v[i];
    tmp >>= skip_bits - (skip_bytes * 8);
    tmp = tmp % ((uint64_t)1 << c);
    return (size_t)tmp;
}

/* arithmetic.rs:51-80 */
typedef struct { int tag; /* 0 None, 1 Affine, 2 Projective */ g1a a; g1j p; } bucket_t;

static void bucket_add_assign(bucket_t* b, const g1a* other) {
    if (b->tag == 0) { b->tag = 1; b->a = *other; }
    else if (b->tag == 1) { b->p = g1a_add(&b->a, other); b->tag = 2; }
    else { b->p = g1j_madd(&b->p, other); }
}
static g1j bucket_add(const bucket_t* b, g1j other) {
    if (b->tag == 0) return other;
    if (b->tag == 1) return g1j_madd(&other, &b->a);
    return g1j_add(&other, &b->p);
}

/* arithmetic.rs:13-101 multiexp_serial */
static void multiexp_serial(const fe* coeffs, const g1a* bases, size_t len, g1j* acc) {
    uint8_t* repr = (uint8_t*)malloc(32 * (len ? len : 1));
    for (size_t i = 0; i < len; i++) fr_to_repr(coeffs[i], repr + 32 * i); /* :14 */
    size_t c;
    if (len < 4) c = 1;
    else if (len < 32) c = 3;
    else c = (size_t)ceil(log((double)(uint32_t)len)); /* :16-22 */
    size_t segments = (256 / c) + 1;               /* :44 */
    size_t nb = ((size_t)1 << c) - 1;
    bucket_t* buckets = (bucket_t*)malloc(sizeof(bucket_t) * nb);
    for (size_t seg = segments; seg-- > 0;) {
        for (size_t k = 0; k < c; k++) *acc = g1j_double(acc); /* :47-49 */
        for (size_t k = 0; k < nb; k++) buckets[k].tag = 0;
        for (size_t i = 0; i < len; i++) { /* :84-89 */
            size_t d = msm_get_at(seg, c, repr + 32 * i);
            if (d != 0) bucket_add_assign(&buckets[d - 1], &bases[i]);
        }
        g1j running = g1j_identity(); /* :95-99 summation by parts */
        for (size_t k = nb; k-- > 0;) {
            running = bucket_add(&buckets[k], running);
            *acc = g1j_add(acc, &running);
        }
    }
    free(buckets);
    free(repr);
}

typedef struct { const fe* coeffs; const g1a* bases; size_t len; g1j acc; } msm_job_t;
static void* msm_worker(void* arg) {
    msm_job_t* j = (msm_job_t*)arg;
    j->acc = g1j_identity();
    multiexp_serial(j->coeffs, j->bases, j->len, &j->acc);
    return NULL;
}

/* arithmetic.rs:132-159 best_multiexp; `num_threads` plays rayon::current_num_threads() */
static g1j best_multiexp(const fe* coeffs, const g1a* bases, size_t len, size_t num_threads) {
    if (num_threads < 1) num_threads = 1;
    if (len > num_threads) {
        size_t chunk = len / num_threads;
        size_t num_chunks = (len + chunk - 1) / chunk; /* coeffs.chunks(chunk).len() */
        msm_job_t* jobs = (msm_job_t*)malloc(sizeof(msm_job_t) * num_chunks);
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * num_chunks);
        for (size_t i = 0; i < num_chunks; i++) {
            size_t off = i * chunk;
            jobs[i].coeffs = coeffs + off;
            jobs[i].bases = bases + off;
            jobs[i].len = (off + chunk <= len) ? chunk : (len - off);
            pthread_create(&th[i], NULL, msm_worker, &jobs[i]);
        }
        g1j acc = g1j_identity();
        for (size_t i = 0; i < num_chunks; i++) {
            pthread_join(th[i], NULL);
            acc = g1j_add(&acc, &jobs[i].acc); /* results.iter().fold(identity, |a,b| a + b) :153 */
        }
        free(jobs);
        free(th);
        return acc;
    }
    g1j acc = g1j_identity();
    multiexp_serial(coeffs, bases, len, &acc);
    return acc;
}

/* ------------------------------------------------------------------------------------------------------------------
 * FFT — halo2_proofs/src/arithmetic.rs:171-274 (G = Fr: group_add/sub/scale = field add/sub/mul, field.rs:69-84)
 * ------------------------------------------------------------------------------------------------------------------ */
static size_t bitreverse(size_t n, size_t l) { /* :172-179 */
    size_t r = 0;
    for (size_t i = 0; i < l; i++) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
static uint32_t log2_floor(size_t num) { /* arithmetic.rs log2_floor */
    uint32_t pow = 0;
    while (((size_t)1 << (pow + 1)) <= num) pow++;
    return pow;
}

typedef struct { fe* a; size_t n; size_t twiddle_chunk; const fe* twiddles; int par_depth; } fft_job_t;
static void recursive_butterfly_arithmetic(fe* a, size_t n, size_t twiddle_chunk, const fe* twiddles, int par_depth);
static void* fft_worker(void* arg) {
    fft_job_t* j = (fft_job_t*)arg;
    recursive_butterfly_arithmetic(j->a, j->n, j->twiddle_chunk, j->twiddles, j->par_depth);
    return NULL;
}
/* arithmetic.rs:237-274; par_depth > 0 forks the left half on a thread (rayon::join) */
static void recursive_butterfly_arithmetic(fe* a, size_t n, size_t twiddle_chunk, const fe* twiddles, int par_depth) {
    if (n == 2) {
        fe t = a[1];
        a[1] = a[0];
        a[0] = fr_add(a[0], t);
        a[1] = fr_sub(a[1], t);
        return;
    }
    fe* left = a;
    fe* right = a + n / 2;
    if (par_depth > 0) {
        fft_job_t job = {left, n / 2, twiddle_chunk * 2, twiddles, par_depth - 1};
        pthread_t th;
        pthread_create(&th, NULL, fft_worker, &job);
        recursive_butterfly_arithmetic(right, n / 2, twiddle_chunk * 2, twiddles, par_depth - 1);
        pthread_join(th, NULL);
    } else {
        recursive_butterfly_arithmetic(left, n / 2, twiddle_chunk * 2, twiddles, 0);
        recursive_butterfly_arithmetic(right, n / 2, twiddle_chunk * 2, twiddles, 0);
    }
    /* case when twiddle factor is one */
    fe t = right[0];
    right[0] = left[0];
    left[0] = fr_add(left[0], t);
    right[0] = fr_sub(right[0], t);
    for (size_t i = 1; i < n / 2; i++) {
        fe tt = fr_mul(right[i], twiddles[i * twiddle_chunk]);
        right[i] = left[i];
        left[i] = fr_add(left[i], tt);
        right[i] = fr_sub(right[i], tt);
    }
}

/* arithmetic.rs:171-234 best_fft; `threads` plays rayon::current_num_threads() */
static void best_fft(fe* a, fe omega, uint32_t log_n, size_t threads) {
    if (threads < 1) threads = 1;
    uint32_t log_threads = log2_floor(threads);
    size_t n = (size_t)1 << log_n;
    for (size_t k = 0; k < n; k++) { /* :186-191 */
        size_t rk = bitreverse(k, log_n);
        if (k < rk) { fe t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    size_t nt = n / 2;
    fe* twiddles = (fe*)malloc(sizeof(fe) * (nt ? nt : 1)); /* :194-200 serial scan */
    fe w = fr_one();
    for (size_t i = 0; i < nt; i++) { twiddles[i] = w; w = fr_mul(w, omega); }
    if (log_n <= log_threads) { /* :202-230 iterative */
        size_t chunk = 2, twiddle_chunk = n / 2;
        for (uint32_t s = 0; s < log_n; s++) {
            for (size_t base = 0; base < n; base += chunk) {
                fe* left = a + base;
                fe* right = a + base + chunk / 2;
                fe t = right[0];
                right[0] = left[0];
                left[0] = fr_add(left[0], t);
                right[0] = fr_sub(right[0], t);
                for (size_t i = 1; i < chunk / 2; i++) {
                    fe tt = fr_mul(right[i], twiddles[i * twiddle_chunk]);
                    right[i] = left[i];
                    left[i] = fr_add(left[i], tt);
                    right[i] = fr_sub(right[i], tt);
                }
            }
            chunk *= 2;
            twiddle_chunk /= 2;
        }
    } else if (n >= 2) {
        recursive_butterfly_arithmetic(a, n, 1, twiddles, (int)log_threads); /* :232 */
    }
    free(twiddles);
}

/* ------------------------------------------------------------------------------------------------------------------
 * EvaluationDomain — halo2_proofs/src/poly/domain.rs
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t n; uint32_t k, extended_k; uint64_t quotient_poly_degree; uint32_t t_len;
    fe omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv, ifft_divisor, extended_ifft_divisor,
        barycentric_weight;
    fe t_evaluations[64]; /* inverted, as stored by the reference after batch_invert; len = 2^(extended_k-k) <= 64 here */
} oracle_domain_t;

static fe fr_pow_u64(fe b, uint64_t e) { uint64_t ee[1] = {e}; return fr_pow_vartime(b, ee, 1); }

/* domain.rs:39-142 EvaluationDomain::new(j, k) */
API int oracle_domain_new(uint32_t j, uint32_t k, oracle_domain_t* d) {
    uint64_t quotient_poly_degree = (uint64_t)(j - 1);
    uint64_t n = (uint64_t)1 << k;
    uint32_t extended_k = k;
    while (((uint64_t)1 << extended_k) < n * quotient_poly_degree) extended_k++;
    if (extended_k - k > 6 || extended_k > FR_S) return -1;
    fe extended_omega = fr_root_of_unity();
    for (uint32_t i = extended_k; i < FR_S; i++) extended_omega = fr_square(extended_omega);
    fe omega = extended_omega;
    for (uint32_t i = k; i < extended_k; i++) omega = fr_square(omega);
    fe g_coset = fr_zeta();
    fe g_coset_inv = fr_square(g_coset);
    uint32_t t_len = 0;
    {
        fe orig = fr_pow_u64(fr_zeta(), n);
        fe step = fr_pow_u64(extended_omega, n);
        fe cur = orig;
        for (;;) {
            d->t_evaluations[t_len++] = cur;
            cur = fr_mul(cur, step);
            if (fr_eq(cur, orig)) break;
            if (t_len >= 64) return -2;
        }
        if (t_len != (1u << (extended_k - k))) return -3;
        for (uint32_t i = 0; i < t_len; i++) d->t_evaluations[i] = fr_sub(d->t_evaluations[i], fr_one());
    }
    /* batch_invert (:118-125) == element-wise inversion, exact arithmetic */
    for (uint32_t i = 0; i < t_len; i++) d->t_evaluations[i] = fr_invert(d->t_evaluations[i]);
    d->ifft_divisor = fr_invert(fr_from_u64((uint64_t)1 << k));
    d->extended_ifft_divisor = fr_invert(fr_from_u64((uint64_t)1 << extended_k));
    d->barycentric_weight = fr_invert(fr_from_u64(n));
    d->extended_omega_inv = fr_invert(extended_omega);
    d->omega_inv = fr_invert(omega);
    d->n = n; d->k = k; d->extended_k = extended_k; d->quotient_poly_degree = quotient_poly_degree; d->t_len = t_len;
    d->omega = omega; d->extended_omega = extended_omega; d->g_coset = g_coset; d->g_coset_inv = g_coset_inv;
    return 0;
}

/* domain.rs:366-374 ifft */
static void domain_ifft(fe* a, fe omega_inv, uint32_t log_n, fe divisor, size_t threads) {
    best_fft(a, omega_inv, log_n, threads);
    size_t n = (size_t)1 << log_n;
    for (size_t i = 0; i < n; i++) a[i] = fr_mul(a[i], divisor);
}
/* domain.rs:347-363 distribute_powers_zeta */
static void distribute_powers_zeta(const oracle_domain_t* d, fe* a, size_t len, int into_coset) {
    fe cp[2];
    if (into_coset) { cp[0] = d->g_coset; cp[1] = d->g_coset_inv; } else { cp[0] = d->g_coset_inv; cp[1] = d->g_coset; }
    for (size_t index = 0; index < len; index++) {
        size_t i = index % 3;
        if (i != 0) a[index] = fr_mul(a[index], cp[i - 1]);
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * exported C API (ctypes-friendly: flat uint64 arrays; fe = 4 limbs, g1a = 8 limbs, g1j = 12 limbs)
 * ------------------------------------------------------------------------------------------------------------------ */

API void oracle_fr_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    fe x, y, r;
    memcpy(&x, a, 32); memcpy(&y, b, 32);
    switch (op) {
        case 0: r = fr_add(x, y); break;   case 1: r = fr_sub(x, y); break;  case 2: r = fr_mul(x, y); break;
        case 3: r = fr_square(x); break;   case 4: r = fr_neg(x); break;     case 5: r = fr_invert(x); break;
        case 6: r = fr_dbl(x); break;      case 7: r = fr_montgomery_reduce_short(x); break;
        case 8: r = fr_from_raw(x.l[0], x.l[1], x.l[2], x.l[3]); break;
        default: r = fr_zero();
    }
    memcpy(out, &r, 32);
}
API void oracle_fq_op(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    fe x, y, r;
    memcpy(&x, a, 32); memcpy(&y, b, 32);
    switch (op) {
        case 0: r = fq_add(x, y); break;   case 1: r = fq_sub(x, y); break;  case 2: r = fq_mul(x, y); break;
        case 3: r = fq_square(x); break;   case 4: r = fq_neg(x); break;     case 5: r = fq_invert(x); break;
        case 6: r = fq_dbl(x); break;      case 7: r = fq_montgomery_reduce_short(x); break;
        case 8: r = fq_from_raw(x.l[0], x.l[1], x.l[2], x.l[3]); break;
        default: r = fq_zero();
    }
    memcpy(out, &r, 32);
}
API void oracle_fr_from_u512(const uint64_t* limbs8, uint64_t* out) { fe r = fr_from_u512(limbs8); memcpy(out, &r, 32); }
API void oracle_fq_from_u512(const uint64_t* limbs8, uint64_t* out) { fe r = fq_from_u512(limbs8); memcpy(out, &r, 32); }
API void oracle_fr_to_repr(const uint64_t* a, uint8_t* out) { fe x; memcpy(&x, a, 32); fr_to_repr(x, out); }
API void oracle_fr_pow(const uint64_t* a, const uint64_t* e4, uint64_t* out) { fe x; memcpy(&x, a, 32); fe r = fr_pow_vartime(x, e4, 4); memcpy(out, &r, 32); }
/* which: 0 ROOT_OF_UNITY 1 ROOT_OF_UNITY_INV 2 TWO_INV 3 DELTA 4 ZETA 5 ONE(R) 6 GENERATOR(7) */
API void oracle_fr_const(int which, uint64_t* out) {
    fe r;
    switch (which) {
        case 0: r = fr_root_of_unity(); break; case 1: r = fr_root_of_unity_inv(); break; case 2: r = fr_two_inv(); break;
        case 3: r = fr_delta(); break;         case 4: r = fr_zeta(); break;             case 5: r = fr_one(); break;
        default: r = fr_from_raw(7, 0, 0, 0);
    }
    memcpy(out, &r, 32);
}

/* G1 ops on flat arrays */
API void oracle_g1_generator(uint64_t* out_aff) { g1a g = g1a_generator(); memcpy(out_aff, &g, 64); }
API int oracle_g1_is_on_curve(const uint64_t* aff) { g1a p; memcpy(&p, aff, 64); return g1a_is_on_curve(&p); }
API void oracle_g1_add_jj(const uint64_t* a, const uint64_t* b, uint64_t* out) { g1j x, y; memcpy(&x, a, 96); memcpy(&y, b, 96); g1j r = g1j_add(&x, &y); memcpy(out, &r, 96); }
API void oracle_g1_add_ja(const uint64_t* a, const uint64_t* b, uint64_t* out) { g1j x; g1a y; memcpy(&x, a, 96); memcpy(&y, b, 64); g1j r = g1j_madd(&x, &y); memcpy(out, &r, 96); }
API void oracle_g1_add_aa(const uint64_t* a, const uint64_t* b, uint64_t* out) { g1a x, y; memcpy(&x, a, 64); memcpy(&y, b, 64); g1j r = g1a_add(&x, &y); memcpy(out, &r, 96); }
API void oracle_g1_double(const uint64_t* a, uint64_t* out) { g1j x; memcpy(&x, a, 96); g1j r = g1j_double(&x); memcpy(out, &r, 96); }
API void oracle_g1_neg_a(const uint64_t* a, uint64_t* out) { g1a x; memcpy(&x, a, 64); g1a r = g1a_neg(&x); memcpy(out, &r, 64); }
API void oracle_g1_to_affine(const uint64_t* a, uint64_t* out) { g1j x; memcpy(&x, a, 96); g1a r = g1j_to_affine(&x); memcpy(out, &r, 64); }
API void oracle_g1_to_curve(const uint64_t* a, uint64_t* out) { g1a x; memcpy(&x, a, 64); g1j r = g1a_to_curve(&x); memcpy(out, &r, 96); }
API void oracle_g1_batch_normalize(const uint64_t* p, uint64_t* q, size_t n) { g1j_batch_normalize((const g1j*)p, (g1a*)q, n); }
API void oracle_g1_mul_a(const uint64_t* a, const uint64_t* s, uint64_t* out) { g1a x; fe k; memcpy(&x, a, 64); memcpy(&k, s, 32); g1j r = g1a_mul(&x, k); memcpy(out, &r, 96); }
API void oracle_g1_mul_j(const uint64_t* a, const uint64_t* s, uint64_t* out) { g1j x; fe k; memcpy(&x, a, 96); memcpy(&k, s, 32); g1j r = g1j_mul(&x, k); memcpy(out, &r, 96); }
API void oracle_g1_to_bytes(const uint64_t* a, uint8_t* out) { g1a x; memcpy(&x, a, 64); g1a_to_bytes(&x, out); }

/* best_multiexp -> Jacobian (96 B) and its affine normal form (64 B) */
API void oracle_best_multiexp(const uint64_t* coeffs, const uint64_t* bases, size_t len, size_t num_threads, uint64_t* out_jac,
                              uint64_t* out_aff) {
    g1j r = best_multiexp((const fe*)coeffs, (const g1a*)bases, len, num_threads);
    if (out_jac) memcpy(out_jac, &r, 96);
    if (out_aff) { g1a a = g1j_to_affine(&r); memcpy(out_aff, &a, 64); }
}
API void oracle_best_fft(uint64_t* a, const uint64_t* omega, uint32_t log_n, size_t threads) {
    fe w; memcpy(&w, omega, 32);
    best_fft((fe*)a, w, log_n, threads);
}
API void oracle_ifft(uint64_t* a, const uint64_t* omega_inv, uint32_t log_n, const uint64_t* divisor, size_t threads) {
    fe w, dv; memcpy(&w, omega_inv, 32); memcpy(&dv, divisor, 32);
    domain_ifft((fe*)a, w, log_n, dv, threads);
}
/* domain.rs:238-248 lagrange_to_coeff (in place) */
API void oracle_lagrange_to_coeff(const oracle_domain_t* d, uint64_t* a, size_t threads) {
    domain_ifft((fe*)a, d->omega_inv, d->k, d->ifft_divisor, threads);
}
/* domain.rs:252-266 coeff_to_extended: in = n coefficients, out = 2^extended_k evaluations on the zeta-coset */
API void oracle_coeff_to_extended(const oracle_domain_t* d, const uint64_t* in, uint64_t* out, size_t threads) {
    size_t n = (size_t)d->n, en = (size_t)1 << d->extended_k;
    fe* o = (fe*)out;
    memcpy(o, in, 32 * n);
    distribute_powers_zeta(d, o, n, 1);
    for (size_t i = n; i < en; i++) o[i] = fr_zero();
    best_fft(o, d->extended_omega, d->extended_k, threads);
}
/* domain.rs:319-338 divide_by_vanishing_poly (in place on 2^extended_k values) */
API void oracle_divide_by_vanishing_poly(const oracle_domain_t* d, uint64_t* a) {
    size_t en = (size_t)1 << d->extended_k;
    fe* h = (fe*)a;
    for (size_t i = 0; i < en; i++) h[i] = fr_mul(h[i], d->t_evaluations[i % d->t_len]);
}
/* domain.rs:293-315 extended_to_coeff: in place; caller reads the first n*quotient_poly_degree values (truncate) */
API size_t oracle_extended_to_coeff(const oracle_domain_t* d, uint64_t* a, size_t threads) {
    size_t en = (size_t)1 << d->extended_k;
    domain_ifft((fe*)a, d->extended_omega_inv, d->extended_k, d->extended_ifft_divisor, threads);
    distribute_powers_zeta(d, (fe*)a, en, 0);
    return (size_t)(d->n * d->quotient_poly_degree);
}

/* kzg/commitment.rs:209-276 ParamsKZG::setup_from_toxic_waste(k, s): g[i] = [s^i]G, g_lagrange[i] = [L_i(s)]G.
 * (The reference walks `current_g *= s` inside rayon chunks; the affine normal forms are identical.) */
API void oracle_params_setup(uint32_t k, const uint64_t* s_, uint64_t* g_out, uint64_t* g_lagrange_out) {
    fe s; memcpy(&s, s_, 32);
    size_t n = (size_t)1 << k;
    g1a g1 = g1a_generator();
    g1j* proj = (g1j*)malloc(sizeof(g1j) * n);
    fe sp = fr_one();
    for (size_t i = 0; i < n; i++) { proj[i] = g1a_mul(&g1, sp); sp = fr_mul(sp, s); }
    g1j_batch_normalize(proj, (g1a*)g_out, n);
    fe root = fr_invert(fr_root_of_unity_inv());
    for (uint32_t i = k; i < FR_S; i++) root = fr_square(root);
    fe n_inv = fr_invert(fr_from_u64((uint64_t)n));
    fe multiplier = fr_mul(fr_sub(fr_pow_u64(s, (uint64_t)n), fr_one()), n_inv);
    for (size_t i = 0; i < n; i++) {
        fe root_pow = fr_pow_u64(root, (uint64_t)i);
        fe scalar = fr_mul(fr_mul(multiplier, root_pow), fr_invert(fr_sub(s, root_pow)));
        proj[i] = g1a_mul(&g1, scalar);
    }
    g1j_batch_normalize(proj, (g1a*)g_lagrange_out, n);
    free(proj);
}

/* kzg/commitment.rs:71-178 TableSRS::setup_from_toxic_waste (G1 parts): g1, g1_lagrange, g_lagrange_opening_at_0 */
API void oracle_table_srs_setup(size_t g1_len, const uint64_t* s_, uint64_t* g1_out, uint64_t* g1_lagrange_out,
                                uint64_t* opening_at_0_out) {
    uint32_t k = log2_floor(g1_len);
    oracle_params_setup(k, s_, g1_out, g1_lagrange_out); /* identical formulas (:86-141 vs :209-262) */
    fe root = fr_invert(fr_root_of_unity_inv());
    for (uint32_t i = k; i < FR_S; i++) root = fr_square(root);
    fe n_inv = fr_invert(fr_from_u64((uint64_t)g1_len));
    const g1a* g1 = (const g1a*)g1_out;
    const g1a* gl = (const g1a*)g1_lagrange_out;
    g1j last_power_scaled = g1a_mul(&g1[g1_len - 1], n_inv); /* :161 */
    g1j neg_last = g1j_neg(&last_power_scaled);
    fe rp = fr_one();
    for (size_t i = 0; i < g1_len; i++) { /* :153-168: l_i * w^{-i} - last_power_scaled */
        fe w_inv_i = fr_invert(rp);
        g1j t = g1a_mul(&gl[i], w_inv_i);
        g1j r = g1j_add(&t, &neg_last);
        g1a ra = g1j_to_affine(&r);
        memcpy(opening_at_0_out + 8 * i, &ra, 64);
        rp = fr_mul(rp, root);
    }
}

/* static_lookup/prover.rs:167-170 (m_cm) and :245-257 (a_cm / qa_cm / a0_cm): serial loop over the sparse support in
 * ascending index order (BTreeMap iteration), `bases[index] * scalar + acc`. */
API void oracle_sparse_commit(const uint64_t* bases, const uint32_t* idx, const uint64_t* scalars, size_t m, uint64_t* out_aff) {
    const g1a* b = (const g1a*)bases;
    g1j acc = g1j_identity();
    for (size_t j = 0; j < m; j++) {
        fe s; memcpy(&s, scalars + 4 * j, 32);
        g1j t = g1a_mul(&b[idx[j]], s);
        acc = g1j_add(&t, &acc);
    }
    g1a a = g1j_to_affine(&acc);
    memcpy(out_aff, &a, 64);
}

/* ------------------------------------------------------------------------------------------------------------------
 * synthetic inputs; the CUDA library has its own generator with the same definition, so tests can
 * also cross-check the generators against each other.
 *   scalars: Fr::random-like (fr.rs:159-170): from_u512 of 8 splitmix64 outputs, counter = 8*i+j
 *   bases  : P_i = [s0 + i*d]G for s0,d = the first two scalars of stream `seed`; normalised in batches
 * ------------------------------------------------------------------------------------------------------------------ */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}
static fe synth_scalar(uint64_t seed, uint64_t i) {
    uint64_t w[8];
    for (int j = 0; j < 8; j++) w[j] = splitmix64(seed * 0x100000001b3ULL + 8 * i + (uint64_t)j);
    return fr_from_u512(w);
}
API void oracle_synth_scalars(uint64_t seed, size_t start, size_t n, uint64_t* out) {
    for (size_t i = 0; i < n; i++) { fe s = synth_scalar(seed, start + i); memcpy(out + 4 * i, &s, 32); }
}
typedef struct { uint64_t seed; size_t start, n; g1a* out; } walk_job_t;
static void* walk_worker(void* arg) {
    walk_job_t* j = (walk_job_t*)arg;
    fe s0 = synth_scalar(j->seed, 0), d = synth_scalar(j->seed, 1);
    g1a g = g1a_generator();
    g1j dj = g1a_mul(&g, d);
    g1a da = g1j_to_affine(&dj);
    fe k = fr_add(s0, fr_mul(d, fr_from_u64((uint64_t)j->start)));
    g1j cur = g1a_mul(&g, k);
    const size_t B = 1024;
    g1j* buf = (g1j*)malloc(sizeof(g1j) * B);
    for (size_t off = 0; off < j->n; off += B) {
        size_t m = (j->n - off < B) ? (j->n - off) : B;
        for (size_t i = 0; i < m; i++) { buf[i] = cur; cur = g1j_madd(&cur, &da); }
        g1j_batch_normalize(buf, j->out + off, m);
    }
    free(buf);
    return NULL;
}
API void oracle_synth_bases(uint64_t seed, size_t n, size_t threads, uint64_t* out) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    walk_job_t jobs[256];
    pthread_t th[256];
    size_t chunk = (n + threads - 1) / threads, used = 0;
    for (size_t t = 0; t < threads; t++) {
        size_t st = t * chunk;
        if (st >= n) break;
        jobs[t].seed = seed; jobs[t].start = st; jobs[t].n = (st + chunk <= n) ? chunk : (n - st); jobs[t].out = (g1a*)out + st;
        pthread_create(&th[t], NULL, walk_worker, &jobs[t]);
        used++;
    }
    for (size_t t = 0; t < used; t++) pthread_join(th[t], NULL);
}
API size_t oracle_domain_sizeof(void) { return sizeof(oracle_domain_t); }
API int oracle_hw_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* ------------------------------------------------------------------------------------------------------------------
 * g_to_lagrange — halo2_proofs/src/arithmetic.rs:277-301: best_fft over G = G1 (group_add / group_sub = Jacobian
 * add / sub, group_scale = 256-bit scalar multiplication), then * n_inv, then batch_normalize. Reached from
 * ParamsKZG::downsize (poly/kzg/commitment.rs:482-490). Serial iterative form of best_fft (arithmetic.rs:202-230).
 * ------------------------------------------------------------------------------------------------------------------ */
static void best_fft_g1(g1j* a, fe omega, uint32_t log_n) {
    size_t n = (size_t)1 << log_n;
    for (size_t k = 0; k < n; k++) {
        size_t rk = bitreverse(k, log_n);
        if (k < rk) { g1j t = a[rk]; a[rk] = a[k]; a[k] = t; }
    }
    size_t nt = n / 2;
    fe* twiddles = (fe*)malloc(sizeof(fe) * (nt ? nt : 1));
    fe w = fr_one();
    for (size_t i = 0; i < nt; i++) { twiddles[i] = w; w = fr_mul(w, omega); }
    size_t chunk = 2, twiddle_chunk = n / 2;
    for (uint32_t s = 0; s < log_n; s++) {
        for (size_t base = 0; base < n; base += chunk) {
            g1j* left = a + base;
            g1j* right = a + base + chunk / 2;
            for (size_t i = 0; i < chunk / 2; i++) {
                g1j t = right[i];
                if (i != 0) t = g1j_mul(&t, twiddles[i * twiddle_chunk]); /* twiddle one is skipped (:213-219) */
                g1j nt_ = g1j_neg(&t);
                right[i] = g1j_add(&left[i], &nt_);
                left[i] = g1j_add(&left[i], &t);
            }
        }
        chunk *= 2;
        twiddle_chunk /= 2;
    }
    free(twiddles);
}

API void oracle_g_to_lagrange(const uint64_t* g_affine, uint32_t k, uint64_t* out_affine) {
    size_t n = (size_t)1 << k;
    uint64_t kk[1] = {k};
    fe n_inv = fr_pow_vartime(fr_two_inv(), kk, 1);      /* TWO_INV.pow_vartime(&[k]) :278 */
    fe omega_inv = fr_root_of_unity_inv();
    for (uint32_t i = k; i < FR_S; i++) omega_inv = fr_square(omega_inv); /* :279-282 */
    g1j* p = (g1j*)malloc(sizeof(g1j) * n);
    for (size_t i = 0; i < n; i++) p[i] = g1a_to_curve(&((const g1a*)g_affine)[i]);
    best_fft_g1(p, omega_inv, k);                        /* :285 */
    for (size_t i = 0; i < n; i++) p[i] = g1j_mul(&p[i], n_inv); /* :286-290 */
    g1j_batch_normalize(p, (g1a*)out_affine, n);         /* :292-298 */
    free(p);
}

/* ------------------------------------------------------------------------------------------------------------------
 * eval_polynomial / kate_division — halo2_proofs/src/arithmetic.rs:304-329, 351-387 (serial forms)
 * ------------------------------------------------------------------------------------------------------------------ */
static fe eval_polynomial(const fe* poly, size_t n, fe point) { /* :305-309 Horner fold from the top coefficient */
    fe acc = fr_zero();
    for (size_t i = n; i-- > 0;) acc = fr_add(fr_mul(acc, point), poly[i]);
    return acc;
}
/* q has n-1 entries: a(X) - a(b) = q(X) (X - b) */
static void kate_division(const fe* a, size_t n, fe b, fe* q) { /* :351-368 */
    fe nb = fr_neg(b);
    fe tmp = fr_zero();
    for (size_t k = n - 1; k-- > 0;) {
        fe lead = fr_sub(a[k + 1], tmp);
        q[k] = lead;
        tmp = fr_mul(lead, nb);
    }
}
API void oracle_eval_polynomial(const uint64_t* poly, size_t n, const uint64_t* point, uint64_t* out) {
    fe p; memcpy(&p, point, 32);
    fe r = eval_polynomial((const fe*)poly, n, p);
    memcpy(out, &r, 32);
}
API void oracle_kate_division(const uint64_t* a, size_t n, const uint64_t* b, uint64_t* q) {
    fe bb; memcpy(&bb, b, 32);
    kate_division((const fe*)a, n, bb, (fe*)q);
}

/* ------------------------------------------------------------------------------------------------------------------
 * StaticTableValues::new — halo2_proofs/src/plonk/static_lookup.rs:77-126: the cached quotient commitments `qs` of the
 * CQ argument: table_coeffs = ifft(values); for every root g_i = w^i: quotient = kate_division(table_coeffs, g_i) scaled
 * by g_i / N; qs[i] = best_multiexp(quotient, srs_g1[..N-1]). O(N^2) ("TODO: THIS SHOULD BE DONE WITH FK METHOD" :107).
 * ------------------------------------------------------------------------------------------------------------------ */
API void oracle_cq_table_qs(const uint64_t* values, size_t size, const uint64_t* srs_g1, size_t threads, uint64_t* qs_affine) {
    uint32_t k = log2_floor(size);
    oracle_domain_t d;
    oracle_domain_new(2, k, &d);
    fe n_inv = fr_invert(fr_from_u64((uint64_t)size));
    fe* coeffs = (fe*)malloc(sizeof(fe) * size);
    memcpy(coeffs, values, sizeof(fe) * size);
    domain_ifft(coeffs, d.omega_inv, k, d.ifft_divisor, threads); /* :99-105 */
    fe* quot = (fe*)malloc(sizeof(fe) * (size > 1 ? size - 1 : 1));
    fe gi = fr_one();
    for (size_t i = 0; i < size; i++) { /* :108-119 */
        g1j acc = g1j_identity();
        if (size > 1) {
            kate_division(coeffs, size, gi, quot);
            fe sc = fr_mul(gi, n_inv);
            for (size_t j = 0; j + 1 < size; j++) quot[j] = fr_mul(quot[j], sc); /* v * g_i * n_inv */
            acc = best_multiexp(quot, (const g1a*)srs_g1, size - 1, threads);
        }
        g1a a = g1j_to_affine(&acc);
        memcpy(qs_affine + 8 * i, &a, 64);
        gi = fr_mul(gi, d.omega);
    }
    free(quot);
    free(coeffs);
}

/* ------------------------------------------------------------------------------------------------------------------
 * evaluate_h pieces — halo2_proofs/src/plonk/evaluation.rs
 *   GraphEvaluator::evaluate (:718-775) over the serialised graph (same word format the CUDA path takes, see
 *   include/cqb200.h): ValueSource = 2 words {kind | rot_idx << 8, index}; kinds follow the enum order (:41-65);
 *   Calculation = {op, target, ...} with op following the enum order (:114-132).
 *   static-lookup (CQ) term (:533-548) and permutation term (:376-452).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
    const fe* constants; const int32_t* rotations; uint32_t n_rot; const uint32_t* code; uint32_t n_calc;
    const fe* const* fixed; const fe* const* advice; const fe* const* instance; const fe* challenges;
    fe beta, gamma, theta, y;
} graph_ctx_t;

static size_t get_rotation_idx(size_t idx, int32_t rot, int32_t rot_scale, int64_t isize) { /* :37-39 rem_euclid */
    int64_t v = ((int64_t)idx + (int64_t)rot * rot_scale) % isize;
    if (v < 0) v += isize;
    return (size_t)v;
}
static fe vs_get(const graph_ctx_t* g, const uint32_t* w, const size_t* rots, const fe* inter, fe prev) { /* :69-109 */
    uint32_t kind = w[0] & 0xff, rot = w[0] >> 8, idx = w[1];
    switch (kind) {
        case 0: return g->constants[idx];
        case 1: return inter[idx];
        case 2: return g->fixed[idx][rots[rot]];
        case 3: return g->advice[idx][rots[rot]];
        case 4: return g->instance[idx][rots[rot]];
        case 5: return g->challenges[idx];
        case 6: return g->beta;
        case 7: return g->gamma;
        case 8: return g->theta;
        case 9: return g->y;
        default: return prev;
    }
}
API void oracle_graph_evaluate(const uint64_t* constants, const int32_t* rotations, uint32_t n_rot, const uint32_t* code, uint32_t n_calc,
                               uint32_t num_intermediates, const uint64_t* const* fixed, const uint64_t* const* advice,
                               const uint64_t* const* instance, const uint64_t* challenges, const uint64_t* bgty /* beta,gamma,theta,y */,
                               uint64_t* values, size_t size, int32_t rot_scale) {
    graph_ctx_t g;
    g.constants = (const fe*)constants; g.rotations = rotations; g.n_rot = n_rot; g.code = code; g.n_calc = n_calc;
    g.fixed = (const fe* const*)fixed; g.advice = (const fe* const*)advice; g.instance = (const fe* const*)instance;
    g.challenges = (const fe*)challenges;
    memcpy(&g.beta, bgty, 32); memcpy(&g.gamma, bgty + 4, 32); memcpy(&g.theta, bgty + 8, 32); memcpy(&g.y, bgty + 12, 32);
    fe* inter = (fe*)malloc(sizeof(fe) * (num_intermediates ? num_intermediates : 1));
    size_t* rots = (size_t*)malloc(sizeof(size_t) * (n_rot ? n_rot : 1));
    fe* vals = (fe*)values;
    for (size_t idx = 0; idx < size; idx++) {
        for (uint32_t r = 0; r < n_rot; r++) rots[r] = get_rotation_idx(idx, rotations[r], rot_scale, (int64_t)size); /* :733-736 */
        fe prev = vals[idx], last = fr_zero();
        const uint32_t* pc = code;
        for (uint32_t c = 0; c < n_calc; c++) { /* :739-757 */
            uint32_t op = pc[0], target = pc[1];
            fe r;
            switch (op) { /* Calculation::evaluate :135-193 */
                case 0: r = fr_add(vs_get(&g, pc + 2, rots, inter, prev), vs_get(&g, pc + 4, rots, inter, prev)); pc += 6; break;
                case 1: r = fr_sub(vs_get(&g, pc + 2, rots, inter, prev), vs_get(&g, pc + 4, rots, inter, prev)); pc += 6; break;
                case 2: r = fr_mul(vs_get(&g, pc + 2, rots, inter, prev), vs_get(&g, pc + 4, rots, inter, prev)); pc += 6; break;
                case 3: r = fr_square(vs_get(&g, pc + 2, rots, inter, prev)); pc += 4; break;
                case 4: r = fr_dbl(vs_get(&g, pc + 2, rots, inter, prev)); pc += 4; break;
                case 5: r = fr_neg(vs_get(&g, pc + 2, rots, inter, prev)); pc += 4; break;
                case 6: { /* Horner(start, parts, factor): words = start, factor, nparts, parts... */
                    fe factor = vs_get(&g, pc + 4, rots, inter, prev);
                    r = vs_get(&g, pc + 2, rots, inter, prev);
                    uint32_t np = pc[6];
                    for (uint32_t p = 0; p < np; p++) r = fr_add(fr_mul(r, factor), vs_get(&g, pc + 7 + 2 * p, rots, inter, prev));
                    pc += 7 + 2 * np;
                    break;
                }
                default: r = vs_get(&g, pc + 2, rots, inter, prev); pc += 4; break; /* Store */
            }
            inter[target] = r;
            last = r;
        }
        vals[idx] = n_calc ? last : fr_zero(); /* :760-765 */
    }
    free(inter);
    free(rots);
}
/* evaluation.rs:533-548: value = value * y + (b_coset * (f_coset * l_active_row + beta) - 1) */
API void oracle_cq_lookup_h(uint64_t* values, const uint64_t* b_coset, const uint64_t* f_coset, const uint64_t* l_active_row,
                            const uint64_t* beta_, const uint64_t* y_, size_t size) {
    fe beta, y; memcpy(&beta, beta_, 32); memcpy(&y, y_, 32);
    fe* v = (fe*)values; const fe* b = (const fe*)b_coset; const fe* f = (const fe*)f_coset; const fe* l = (const fe*)l_active_row;
    for (size_t i = 0; i < size; i++)
        v[i] = fr_add(fr_mul(v[i], y), fr_sub(fr_mul(b[i], fr_add(fr_mul(f[i], l[i]), beta)), fr_one()));
}
/* evaluation.rs:376-452 permutation constraints. sets: nsets product cosets; columns / perm cosets: ncols each, chunked by
 * chunk_len per set */
API void oracle_permutation_h(uint64_t* values, size_t size, int32_t rot_scale, int32_t last_rotation, uint32_t chunk_len,
                              const uint64_t* const* sets, uint32_t nsets, const uint64_t* const* columns,
                              const uint64_t* const* perm_cosets, uint32_t ncols, const uint64_t* l0_, const uint64_t* l_last_,
                              const uint64_t* l_active_, const uint64_t* beta_, const uint64_t* gamma_, const uint64_t* y_,
                              const uint64_t* extended_omega_) {
    fe beta, gamma, y, ew; memcpy(&beta, beta_, 32); memcpy(&gamma, gamma_, 32); memcpy(&y, y_, 32); memcpy(&ew, extended_omega_, 32);
    fe* v = (fe*)values; const fe* l0 = (const fe*)l0_; const fe* l_last = (const fe*)l_last_; const fe* l_active = (const fe*)l_active_;
    fe one = fr_one(), delta = fr_delta();
    fe delta_start = fr_mul(beta, fr_zeta()); /* :383 */
    fe beta_term = fr_one();                  /* extended_omega^start, start = 0 (:390) */
    for (size_t idx = 0; idx < size; idx++) {
        size_t r_next = get_rotation_idx(idx, 1, rot_scale, (int64_t)size);
        size_t r_last = get_rotation_idx(idx, last_rotation, rot_scale, (int64_t)size);
        const fe* first = (const fe*)sets[0]; const fe* last = (const fe*)sets[nsets - 1];
        v[idx] = fr_add(fr_mul(v[idx], y), fr_mul(fr_sub(one, first[idx]), l0[idx]));                                   /* :397-399 */
        v[idx] = fr_add(fr_mul(v[idx], y), fr_mul(fr_sub(fr_mul(last[idx], last[idx]), last[idx]), l_last[idx]));       /* :402-406 */
        for (uint32_t s = 1; s < nsets; s++)                                                                             /* :409-417 */
            v[idx] = fr_add(fr_mul(v[idx], y), fr_mul(fr_sub(((const fe*)sets[s])[idx], ((const fe*)sets[s - 1])[r_last]), l0[idx]));
        fe current_delta = fr_mul(delta_start, beta_term);                                                               /* :423 */
        for (uint32_t s = 0; s < nsets; s++) {                                                                           /* :424-449 */
            uint32_t c0 = s * chunk_len, c1 = c0 + chunk_len < ncols ? c0 + chunk_len : ncols;
            fe left = ((const fe*)sets[s])[r_next];
            for (uint32_t c = c0; c < c1; c++)
                left = fr_mul(left, fr_add(fr_add(((const fe*)columns[c])[idx], fr_mul(beta, ((const fe*)perm_cosets[c])[idx])), gamma));
            fe right = ((const fe*)sets[s])[idx];
            for (uint32_t c = c0; c < c1; c++) {
                right = fr_mul(right, fr_add(fr_add(((const fe*)columns[c])[idx], current_delta), gamma));
                current_delta = fr_mul(current_delta, delta);
            }
            v[idx] = fr_add(fr_mul(v[idx], y), fr_mul(fr_sub(left, right), l_active[idx]));
        }
        beta_term = fr_mul(beta_term, ew);                                                                               /* :450 */
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * permutation::Argument::commit — halo2_proofs/src/plonk/permutation/prover.rs:82-166, ONE column set (one iteration of
 * the `for (columns, permutations) in self.columns.chunks(chunk_len)...` loop): the grand-product vector z in Lagrange
 * form BEFORE the blinding rows are overwritten (:152-155 draws them from the rng; the caller owns that).
 *   modified_values[i]  = prod_j (beta * perm_j[i] + gamma + col_j[i])            (:103-118)
 *   modified_values    <- 1 / modified_values  (batch_invert == element-wise inverse, exact)   (:121)
 *   modified_values[i] *= prod_j (deltaomega_j * omega^i * beta + gamma + col_j[i]),  deltaomega_{j+1} = deltaomega_j * DELTA  (:125-144)
 *   z[0] = last_z ; z[row] = z[row-1] * modified_values[row-1]                     (:157-163)
 * deltaomega_io: in = DELTA^(index of the set's first column), out = value for the next set (:144).
 * ------------------------------------------------------------------------------------------------------------------ */
API void oracle_permutation_product(const uint64_t* const* columns, const uint64_t* const* perms, uint32_t ncols, size_t n,
                                    const uint64_t* beta_, const uint64_t* gamma_, const uint64_t* omega_, uint64_t* deltaomega_io,
                                    const uint64_t* last_z_, uint64_t* z_out) {
    fe beta, gamma, omega, deltaomega, last_z;
    memcpy(&beta, beta_, 32); memcpy(&gamma, gamma_, 32); memcpy(&omega, omega_, 32); memcpy(&deltaomega, deltaomega_io, 32); memcpy(&last_z, last_z_, 32);
    fe* mv = (fe*)malloc(sizeof(fe) * n);
    for (size_t i = 0; i < n; i++) mv[i] = fr_one();
    for (uint32_t j = 0; j < ncols; j++)
        for (size_t i = 0; i < n; i++)
            mv[i] = fr_mul(mv[i], fr_add(fr_add(fr_mul(beta, ((const fe*)perms[j])[i]), gamma), ((const fe*)columns[j])[i]));
    for (size_t i = 0; i < n; i++) mv[i] = fr_invert(mv[i]);
    for (uint32_t j = 0; j < ncols; j++) {
        fe dw = deltaomega; /* start = 0: deltaomega * omega^0 */
        for (size_t i = 0; i < n; i++) {
            mv[i] = fr_mul(mv[i], fr_add(fr_add(fr_mul(dw, beta), gamma), ((const fe*)columns[j])[i]));
            dw = fr_mul(dw, omega);
        }
        deltaomega = fr_mul(deltaomega, fr_delta());
    }
    fe* z = (fe*)z_out;
    z[0] = last_z;
    for (size_t row = 1; row < n; row++) z[row] = fr_mul(z[row - 1], mv[row - 1]);
    memcpy(deltaomega_io, &deltaomega, 32);
    free(mv);
}

/* ------------------------------------------------------------------------------------------------------------------
 * lookup::prover::Permuted::commit_product — halo2_proofs/src/plonk/lookup/prover.rs:173-262: the plookup grand product in
 * Lagrange form BEFORE the blinding rows are appended (:259 draws them from the rng; the caller owns that): z has n entries,
 * z[0] = 1, z[i] = prod_{r<i} (a_r + beta)(s_r + gamma) / ((a'_r + beta)(s'_r + gamma)).
 * ------------------------------------------------------------------------------------------------------------------ */
API void oracle_lookup_product(const uint64_t* cin, const uint64_t* ctab, const uint64_t* pin, const uint64_t* ptab, size_t n,
                               const uint64_t* beta_, const uint64_t* gamma_, uint64_t* z_out) {
    fe beta, gamma; memcpy(&beta, beta_, 32); memcpy(&gamma, gamma_, 32);
    const fe *a = (const fe*)cin, *s = (const fe*)ctab, *ap = (const fe*)pin, *sp = (const fe*)ptab;
    fe* z = (fe*)z_out;
    fe state = fr_one();
    z[0] = state; /* iter::once(one) scanned from state = one (:249-254) */
    for (size_t i = 0; i + 1 < n; i++) {
        fe lp = fr_mul(fr_add(beta, ap[i]), fr_add(gamma, sp[i]));   /* :214 */
        lp = fr_invert(lp);                                           /* :220 */
        lp = fr_mul(lp, fr_add(a[i], beta));                          /* :229 */
        lp = fr_mul(lp, fr_add(s[i], gamma));                         /* :230 */
        state = fr_mul(state, lp);
        z[i + 1] = state;
    }
}

/* evaluation.rs:458-531: the plookup constraints of ONE lookup folded into values with y. table_value: the lookup
 * GraphEvaluator's output per row (evaluated by the caller with oracle_graph_evaluate). */
API void oracle_lookup_h(uint64_t* values, size_t size, int32_t rot_scale, const uint64_t* table_value_, const uint64_t* product_,
                         const uint64_t* pin_, const uint64_t* ptab_, const uint64_t* l0_, const uint64_t* l_last_, const uint64_t* l_active_,
                         const uint64_t* beta_, const uint64_t* gamma_, const uint64_t* y_) {
    fe beta, gamma, y; memcpy(&beta, beta_, 32); memcpy(&gamma, gamma_, 32); memcpy(&y, y_, 32);
    fe* v = (fe*)values;
    const fe *tv = (const fe*)table_value_, *z = (const fe*)product_, *a = (const fe*)pin_, *s = (const fe*)ptab_;
    const fe *l0 = (const fe*)l0_, *l_last = (const fe*)l_last_, *l_active = (const fe*)l_active_;
    fe one = fr_one();
    for (size_t idx = 0; idx < size; idx++) {
        size_t r_next = get_rotation_idx(idx, 1, rot_scale, (int64_t)size);
        size_t r_prev = get_rotation_idx(idx, -1, rot_scale, (int64_t)size);
        fe a_minus_s = fr_sub(a[idx], s[idx]);
        v[idx] = fr_add(fr_mul(v[idx], y), fr_mul(fr_sub(one, z[idx]), l0[idx]));                                   /* :494-495 */
        v[idx] = fr_add(fr_mul(v[idx], y), fr_mul(fr_sub(fr_mul(z[idx], z[idx]), z[idx]), l_last[idx]));             /* :497-500 */
        fe left = fr_mul(fr_mul(z[r_next], fr_add(a[idx], beta)), fr_add(s[idx], gamma));
        v[idx] = fr_add(fr_mul(v[idx], y), fr_mul(fr_sub(left, fr_mul(z[idx], tv[idx])), l_active[idx]));            /* :506-512 */
        v[idx] = fr_add(fr_mul(v[idx], y), fr_mul(a_minus_s, l0[idx]));                                              /* :516 */
        v[idx] = fr_add(fr_mul(v[idx], y), fr_mul(fr_mul(a_minus_s, fr_sub(a[idx], a[r_prev])), l_active[idx]));     /* :521-525 */
    }
}

/* ---- G2 (keygen-time pieces of the path: [s]G2, the table SRS's G2 powers, the G2 table commitment) --------------------------- */
API void oracle_g2_generator(uint64_t* out_aff) { g2a g = g2a_generator(); memcpy(out_aff, &g, 128); }
API int oracle_g2_is_on_curve(const uint64_t* a) { g2a x; memcpy(&x, a, 128); return g2a_is_on_curve(&x); }
API void oracle_g2_mul_a(const uint64_t* a, const uint64_t* s, uint64_t* out_aff) {
    g2a x; fe k; memcpy(&x, a, 128); memcpy(&k, s, 32);
    g2j r = g2a_mul(&x, k); g2a o = g2j_to_affine(&r); memcpy(out_aff, &o, 128);
}
API void oracle_g2_add_aa(const uint64_t* a, const uint64_t* b, uint64_t* out_aff) {
    g2a x, y; memcpy(&x, a, 128); memcpy(&y, b, 128);
    g2j xj = g2a_to_curve(&x); g2j r = g2j_madd(&xj, &y); g2a o = g2j_to_affine(&r); memcpy(out_aff, &o, 128);
}
API void oracle_g2_neg_a(const uint64_t* a, uint64_t* out_aff) { g2a x; memcpy(&x, a, 128); g2a r = g2a_neg(&x); memcpy(out_aff, &r, 128); }
/* poly/kzg/commitment.rs:94-104 (TableSRS) / :265 (ParamsKZG s_g2): out[i] = [s^i] G2, i < count, affine */
API void oracle_g2_powers(const uint64_t* s_, size_t count, uint64_t* out_aff) {
    fe s; memcpy(&s, s_, 32);
    g2a gen = g2a_generator();
    fe cur = fr_one();
    for (size_t i = 0; i < count; i++) {
        g2j r = g2a_mul(&gen, cur);
        g2a o = g2j_to_affine(&r);
        memcpy(out_aff + 16 * i, &o, 128);
        cur = fr_mul(cur, s);
    }
}
/* best_multiexp::<G2Affine> (plonk/static_lookup.rs:146): only the affine normal form of the sum is canonical, so the oracle adds
 * the scalar multiples one by one */
API void oracle_g2_msm(const uint64_t* bases_aff, const uint64_t* scalars, size_t n, uint64_t* out_aff) {
    g2j acc = g2j_identity();
    for (size_t i = 0; i < n; i++) {
        g2a b; fe k; memcpy(&b, bases_aff + 16 * i, 128); memcpy(&k, scalars + 4 * i, 32);
        g2j t = g2a_mul(&b, k);
        acc = g2j_add(&acc, &t);
    }
    g2a o = g2j_to_affine(&acc);
    memcpy(out_aff, &o, 128);
}

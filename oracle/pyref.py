"""pyref.py — TEST INFRASTRUCTURE. Independent Python big-integer model of the bn256 fields / G1 / MSM / NTT.

It shares no code or algorithm structure with oracle/bn254_oracle.c (textbook affine formulas, pow(x,-1,p), O(n^2) or
plain recursive DFT), so it guards the oracle itself on small sizes. Values cross the boundary as
Montgomery-form integers (x*2^256 mod p), little-endian 4x64 limbs, exactly the reference's in-memory layout
(arithmetic/curves/src/bn256/fr.rs:22-25).
"""
import numpy as np

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  # bn256/fr.rs:16
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  # bn256/fq.rs:18
MONT = 1 << 256
FR_S = 28
FR_GENERATOR = 7
ROOT_OF_UNITY = 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c  # fr.rs:77-82 (canonical)
ZETA = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23           # fr.rs:112-117
G1_GEN = (1, 2)  # bn256/curve.rs:66-67
G1_B = 3


def to_mont(x, p):
    return (x * MONT) % p


def from_mont(x, p):
    return (x * pow(MONT, -1, p)) % p


def int_to_limbs(x):
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def limbs_to_int(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    return sum(int(a[i]) << (64 * i) for i in range(4))


def fr_array_from_ints(vals):
    """canonical ints -> (n,4) uint64 Montgomery limbs"""
    return np.stack([int_to_limbs(to_mont(v % R_MOD, R_MOD)) for v in vals]) if len(vals) else np.zeros((0, 4), np.uint64)


def fr_array_to_ints(arr):
    return [from_mont(limbs_to_int(r), R_MOD) for r in np.asarray(arr, dtype=np.uint64).reshape(-1, 4)]


def g1_affine_from_ints(pts):
    """list of (x,y) canonical ints or None (identity) -> (n,8) uint64 Montgomery limbs; identity = zeros"""
    out = np.zeros((len(pts), 8), np.uint64)
    for i, p in enumerate(pts):
        if p is not None:
            out[i, :4] = int_to_limbs(to_mont(p[0], Q_MOD))
            out[i, 4:] = int_to_limbs(to_mont(p[1], Q_MOD))
    return out


def g1_affine_to_ints(arr):
    res = []
    for r in np.asarray(arr, dtype=np.uint64).reshape(-1, 8):
        x, y = limbs_to_int(r[:4]), limbs_to_int(r[4:])
        res.append(None if (x == 0 and y == 0) else (from_mont(x, Q_MOD), from_mont(y, Q_MOD)))
    return res


# ---- textbook short-Weierstrass arithmetic over Fq (a = 0) ----
def g1_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    y3 = (lam * (x1 - x3) - y1) % Q_MOD
    return (x3, y3)


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % Q_MOD)


def g1_mul(P, k):
    k %= R_MOD
    acc = None
    add = P
    while k:
        if k & 1:
            acc = g1_add(acc, add)
        add = g1_add(add, add)
        k >>= 1
    return acc


def g1_on_curve(P):
    return P is None or (P[1] * P[1] - P[0] ** 3 - G1_B) % Q_MOD == 0


def msm(scalars, points):
    acc = None
    for s, P in zip(scalars, points):
        acc = g1_add(acc, g1_mul(P, s))
    return acc


def g1_compress(P):
    """derive/curve.rs:635-646"""
    if P is None:
        return bytes(32)
    b = bytearray(P[0].to_bytes(32, "little"))
    b[31] |= (P[1] & 1) << 7
    return bytes(b)


def omega_for(k):
    """domain.rs:54-61: ROOT_OF_UNITY^(2^(S-k))"""
    w = ROOT_OF_UNITY
    for _ in range(k, FR_S):
        w = w * w % R_MOD
    return w


def dft(a, omega):
    """out[k] = sum_j a[j] omega^(jk)  (definition of best_fft's result, arithmetic.rs:161-170 doc comment)"""
    n = len(a)
    if n == 1:
        return list(a)
    if n <= 8:
        return [sum(a[j] * pow(omega, j * k, R_MOD) for j in range(n)) % R_MOD for k in range(n)]
    even = dft(a[0::2], omega * omega % R_MOD)
    odd = dft(a[1::2], omega * omega % R_MOD)
    out = [0] * n
    w = 1
    for k in range(n // 2):
        t = w * odd[k] % R_MOD
        out[k] = (even[k] + t) % R_MOD
        out[k + n // 2] = (even[k] - t) % R_MOD
        w = w * omega % R_MOD
    return out


# ---- Fq2 = Fq[u]/(u^2 + 1) and G2: an independent big-integer model (tuples (c0, c1); affine points ((x0,x1),(y0,y1)) or None) ----
G2_GEN = ((0x1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed, 0x198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2),
          (0x12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa, 0x090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b))


def fq2_add(a, b):
    return ((a[0] + b[0]) % Q_MOD, (a[1] + b[1]) % Q_MOD)


def fq2_sub(a, b):
    return ((a[0] - b[0]) % Q_MOD, (a[1] - b[1]) % Q_MOD)


def fq2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q_MOD, (a[0] * b[1] + a[1] * b[0]) % Q_MOD)


def fq2_inv(a):
    t = pow(a[0] * a[0] + a[1] * a[1], -1, Q_MOD)
    return (a[0] * t % Q_MOD, -a[1] * t % Q_MOD)


G2_B = fq2_mul((3, 0), fq2_inv((9, 1)))  # 3 / (9 + u), bn256/curve.rs:85-98


def g2_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    (x1, y1), (x2, y2) = P, Q
    if x1 == x2:
        if fq2_add(y1, y2) == (0, 0):
            return None
        lam = fq2_mul(fq2_mul((3, 0), fq2_mul(x1, x1)), fq2_inv(fq2_add(y1, y1)))
    else:
        lam = fq2_mul(fq2_sub(y2, y1), fq2_inv(fq2_sub(x2, x1)))
    x3 = fq2_sub(fq2_sub(fq2_mul(lam, lam), x1), x2)
    y3 = fq2_sub(fq2_mul(lam, fq2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(P, k):
    k %= R_MOD
    acc, add = None, P
    while k:
        if k & 1:
            acc = g2_add(acc, add)
        add = g2_add(add, add)
        k >>= 1
    return acc


def g2_on_curve(P):
    if P is None:
        return True
    x, y = P
    return fq2_mul(y, y) == fq2_add(fq2_mul(fq2_mul(x, x), x), G2_B)


def g2_affine_to_ints(limbs16):
    a = np.asarray(limbs16, dtype=np.uint64).reshape(4, 4)
    v = [from_mont(limbs_to_int(r), Q_MOD) for r in a]
    if not any(limbs_to_int(r) for r in a):
        return None
    return ((v[0], v[1]), (v[2], v[3]))

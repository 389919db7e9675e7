"""The device field / curve code (csrc/fp.cuh, csrc/ec.cuh) also compiles for the host with an emulated carry flag, so
the exact limb algorithm (even/odd interleaved CIOS, XYZZ formulas and their exceptional branches) is checked here on CPU
against Python big integers — no GPU needed."""
import os
import random
import subprocess

import pytest

from oracle import pyref as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path_factory, name):
    out = tmp_path_factory.mktemp("hostcheck") / name
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(out), os.path.join(ROOT, "tools", name + ".cpp")])
    return str(out)


@pytest.fixture(scope="module")
def fp_bin(tmp_path_factory):
    return _build(tmp_path_factory, "fp_host_check")


@pytest.fixture(scope="module")
def ec_bin(tmp_path_factory):
    return _build(tmp_path_factory, "ec_host_check")


def test_fp_limb_algorithm(fp_bin):
    rng = random.Random(1)
    lines, exp = [], []
    for name, p in (("fr", P.R_MOD), ("fq", P.Q_MOD)):
        rinv = pow(P.MONT, -1, p)
        vals = [0, 1, 2, p - 1, p - 2, P.MONT % p, (1 << 254) % p, (p - 1) // 2]
        vals += [(1 << k) - 1 for k in (32, 64, 96, 128, 160, 192, 224, 253)] + [((1 << 253) - 1) ^ ((1 << k) - 1) for k in (31, 97, 200)]
        vals += [rng.randrange(p) for _ in range(300)]
        for _ in range(500):
            a, b = rng.choice(vals), rng.choice(vals)
            for op, res in (("mul", a * b * rinv % p), ("mulk", a * b * rinv % p), ("add", (a + b) % p), ("sub", (a - b) % p), ("neg", (-a) % p),
                            ("dbl", 2 * a % p), ("sqr", a * a * rinv % p), ("frommont", a * rinv % p), ("tomont", a * P.MONT % p)):
                lines.append(f"{name} {op} {a:064x} {b:064x}")
                exp.append(res)
        for _ in range(400):  # a*b + c*d with a single reduction; operands up to p itself (a negated zero)
            a, b, c, d = (rng.choice(vals + [p]) for _ in range(4))
            lines.append(f"{name} mul2 {a:064x} {b:064x} {c:064x} {d:064x}")
            exp.append((a * b + c * d) * rinv % p)
        # from_u512's first operand is an arbitrary 256-bit value (derive/field.rs:29-48): unreduced multiplicand
        for _ in range(200):
            a, b = rng.randrange(1 << 256), rng.choice(vals)
            for op in ("mul", "mulk"):
                lines.append(f"{name} {op} {a:064x} {b:064x}")
                exp.append(a * b * rinv % p)
        for a in vals[:24]:
            for op in ("inv", "invb"):  # Fermat and binary-Euclid inversions agree (the inverse is unique)
                lines.append(f"{name} {op} {a:064x} {0:064x}")
                exp.append(0 if a == 0 else pow(a * rinv % p, -1, p) * P.MONT % p)
        for a in vals[24:120]:
            lines.append(f"{name} invb {a:064x} {0:064x}")
            exp.append(pow(a * rinv % p, -1, p) * P.MONT % p)
        # the branch-free safegcd inversion (batched-affine bucket accumulation): structured values and 2000 random ones
        for a in vals + [rng.randrange(p) for _ in range(2000)] + [rng.randrange(1 << k) for k in (1, 2, 31, 33, 60, 61, 90, 120, 150, 250) for _ in range(20)]:
            a %= p
            lines.append(f"{name} invs {a:064x} {0:064x}")
            exp.append(0 if a == 0 else pow(a * rinv % p, -1, p) * P.MONT % p)
    out = subprocess.run([fp_bin], input="\n".join(lines), capture_output=True, text=True, check=True).stdout.split()
    assert len(out) == len(exp)
    bad = [(l, o) for l, o, e in zip(lines, out, exp) if int(o, 16) != e]
    assert not bad, bad[:3]


def _m(v):
    return f"{P.to_mont(v, P.Q_MOD):064x}"


def test_xyzz_formulas_and_exceptional_cases(ec_bin):
    rng = random.Random(2)
    pts = [P.g1_mul(P.G1_GEN, rng.randrange(1, P.R_MOD)) for _ in range(8)]
    script, exp = [], []

    def madd(pt):
        script.append(f"madd {_m(pt[0])} {_m(pt[1])}")

    def out(expected):
        script.append("out")
        exp.append(expected)

    # plain chain
    acc = None
    script.append("reset")
    for pt in pts:
        madd(pt)
        acc = P.g1_add(acc, pt)
        out(acc)
    # P + P via madd (doubling branch), then + (-2P) (opposite branch -> identity), then continue from identity
    script.append("reset")
    madd(pts[0]); madd(pts[0]); out(P.g1_add(pts[0], pts[0]))
    m2 = P.g1_neg(P.g1_add(pts[0], pts[0]))
    madd(m2); out(None)
    madd(pts[1]); out(pts[1])
    # full add: acc + saved, with equal and opposite operands
    script.append("reset")
    madd(pts[2]); madd(pts[3]); script.append("save")
    s = P.g1_add(pts[2], pts[3])
    script.append("addsaved"); out(P.g1_add(s, s))          # add(P, P) -> double branch
    script.append("dbl"); out(P.g1_mul(s, 4))
    script.append("reset"); madd(P.g1_neg(pts[2])); madd(P.g1_neg(pts[3])); script.append("addsaved"); out(None)  # P + (-P)
    script.append("addsaved"); out(s)                        # identity + saved
    script.append("reset"); madd(pts[4]); script.append("save"); script.append("reset"); script.append("save")
    madd(pts[5]); script.append("addsaved"); out(pts[5])     # acc + identity
    res = subprocess.run([ec_bin], input="\n".join(script), capture_output=True, text=True, check=True).stdout.strip().split("\n")
    assert len(res) == len(exp)
    rinv = pow(P.MONT, -1, P.Q_MOD)
    for line, e in zip(res, exp):
        x, y = (int(v, 16) for v in line.split())
        got = None if (x == 0 and y == 0) else (x * rinv % P.Q_MOD, y * rinv % P.Q_MOD)
        assert got == e

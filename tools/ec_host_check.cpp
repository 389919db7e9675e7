// Host-side checker for csrc/ec.cuh (XYZZ formulas incl. exceptional cases), carry flag emulated. Commands on stdin:
//   reset | madd <x> <y> | save | addsaved | dbl | out      (x,y: 64 hex digits, Montgomery form)
// `out` prints the affine normal form "x y" (zeros for the identity). Driven by tests/test_fp_host.py.
#include <cstdio>
#include <iostream>
#include <string>
#include "../sha2-on-cq-halo2_b200/csrc/ec.cuh"
using namespace cqb;
static Fq parse(const std::string& s) {
    Fq r;
    for (int i = 0; i < 8; i++) r.l[i] = (uint32_t)strtoul(s.substr(64 - 8 * (i + 1), 8).c_str(), nullptr, 16);
    return r;
}
static void put(const Fq& a) { for (int i = 7; i >= 0; i--) printf("%08x", a.l[i]); }
int main() {
    G1Xyzz acc = G1Xyzz::identity(), saved = G1Xyzz::identity();
    std::string cmd, a, b;
    while (std::cin >> cmd) {
        if (cmd == "reset") acc = G1Xyzz::identity();
        else if (cmd == "madd") { std::cin >> a >> b; Fq x = parse(a), y = parse(b); if (!(x.is_zero() && y.is_zero())) g1_madd(acc, x, y); }
        else if (cmd == "save") saved = acc;
        else if (cmd == "addsaved") g1_add(acc, saved);
        else if (cmd == "dbl") acc = g1_double(acc);
        else if (cmd == "out") { G1Affine p = g1_to_affine(acc); put(p.x); printf(" "); put(p.y); printf("\n"); }
    }
    return 0;
}

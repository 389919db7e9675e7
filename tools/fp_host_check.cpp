// Host-side limb-algorithm checker for csrc/fp.cuh (carry flag emulated). Reads lines "F op a b" (hex, 64 digits,
// big-endian integer strings of the raw limb values) and prints the raw result limbs as hex. Driven by tests/test_fp_host.py.
#include <cstdio>
#include <cstring>
#include <string>
#include <iostream>
#include "../sha2-on-cq-halo2_b200/csrc/fp.cuh"
using namespace cqb;
template <class F> static F parse(const std::string& s) {
    F r;
    for (int i = 0; i < 8; i++) r.l[i] = (uint32_t)strtoul(s.substr(64 - 8 * (i + 1), 8).c_str(), nullptr, 16);
    return r;
}
template <class F> static void put(const F& a) {
    for (int i = 7; i >= 0; i--) printf("%08x", a.l[i]);
    printf("\n");
}
template <class P> static void run(const std::string& op, const std::string& sa, const std::string& sb) {
    typedef Fp<P> F;
    F a = parse<F>(sa), b = parse<F>(sb);
    if (op == "mul") put(fp_mul<P>(a, b));
    else if (op == "mulk") put(fp_mul_kar<P>(a, b));
    else if (op == "sqr") put(fp_sqr<P>(a));
    else if (op == "add") put(fp_add<P>(a, b));
    else if (op == "sub") put(fp_sub<P>(a, b));
    else if (op == "neg") put(fp_neg<P>(a));
    else if (op == "dbl") put(fp_dbl<P>(a));
    else if (op == "inv") put(fp_inv<P>(a));
    else if (op == "invb") put(fp_inv_binary<P>(a));
    else if (op == "invs") put(fp_inv_safegcd<P>(a));
    else if (op == "frommont") put(fp_from_mont<P>(a));
    else if (op == "tomont") put(fp_to_mont<P>(a));
    else printf("?\n");
}
int main() {
    std::string f, op, a, b;
    while (std::cin >> f >> op >> a >> b) {
        if (op == "mul2") {  // a*b + c*d with one reduction
            std::string c, d;
            std::cin >> c >> d;
            if (f == "fr") put(fp_mul2<FrP>(parse<Fr>(a), parse<Fr>(b), parse<Fr>(c), parse<Fr>(d)));
            else put(fp_mul2<FqP>(parse<Fq>(a), parse<Fq>(b), parse<Fq>(c), parse<Fq>(d)));
            continue;
        }
        if (f == "fr") run<FrP>(op, a, b); else run<FqP>(op, a, b);
    }
    return 0;
}

// The C++ create_proof (csrc/host/halo2_b200_prover.hpp) on a circuit handed over by tests/test_gpu_cpp_host.py in a flat binary file:
// the proof bytes it writes must equal the Python mirror's (which tests/test_gpu_full_proof.py ties to the oracle prover and verifies).
// The Blake2b transcript below is test infrastructure: BLAKE2b (RFC 7693) with the reference's personalisation and prefixes
// (halo2_proofs/src/transcript.rs:14-20, 199-240, 297-315).
// usage: test_create_proof <circuit.bin> <proof.out>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "../../sha2-on-cq-halo2_b200/csrc/host/halo2_b200_prover.hpp"

using namespace halo2_b200;

// ---- BLAKE2b-512, unkeyed, with a 16-byte personalisation ----
static const uint64_t B2_IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                  0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static const uint8_t B2_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static inline uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }
static void b2_compress(uint64_t h[8], const uint8_t block[128], uint64_t t, bool last) {
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) memcpy(&m[i], block + 8 * i, 8);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2_IV[i]; }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
    auto G = [&](int r, int i, int a, int b, int c, int d) {
        v[a] = v[a] + v[b] + m[B2_SIGMA[r][2 * i]];     v[d] = rotr64(v[d] ^ v[a], 32);
        v[c] = v[c] + v[d];                             v[b] = rotr64(v[b] ^ v[c], 24);
        v[a] = v[a] + v[b] + m[B2_SIGMA[r][2 * i + 1]]; v[d] = rotr64(v[d] ^ v[a], 16);
        v[c] = v[c] + v[d];                             v[b] = rotr64(v[b] ^ v[c], 63);
    };
    for (int r = 0; r < 12; r++) {
        G(r, 0, 0, 4, 8, 12); G(r, 1, 1, 5, 9, 13); G(r, 2, 2, 6, 10, 14); G(r, 3, 3, 7, 11, 15);
        G(r, 4, 0, 5, 10, 15); G(r, 5, 1, 6, 11, 12); G(r, 6, 2, 7, 8, 13); G(r, 7, 3, 4, 9, 14);
    }
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
static void blake2b_512_personal(const std::vector<uint8_t>& msg, const char personal[16], uint8_t out[64]) {
    uint8_t param[64] = {0};
    param[0] = 64; param[2] = 1; param[3] = 1;  // digest length, fanout, depth
    memcpy(param + 48, personal, 16);
    uint64_t h[8];
    for (int i = 0; i < 8; i++) { uint64_t p; memcpy(&p, param + 8 * i, 8); h[i] = B2_IV[i] ^ p; }
    size_t off = 0;
    uint8_t block[128];
    while (msg.size() - off > 128) { b2_compress(h, msg.data() + off, off + 128, false); off += 128; }
    memset(block, 0, 128);
    memcpy(block, msg.data() + off, msg.size() - off);
    b2_compress(h, block, msg.size(), true);
    memcpy(out, h, 64);
}

// ---- the reference's Blake2bWrite + Challenge255 (every squeeze hashes the absorbed bytes again: state.clone().finalize()) ----
struct Blake2bTranscript : plonk::Transcript {
    std::vector<uint8_t> absorbed, proof;
    static void canonical(const uint64_t mont[4], bool fq, uint8_t out[32]) {
        if (fq) {
            cqb::Fq a;
            for (int i = 0; i < 4; i++) { a.l[2 * i] = (uint32_t)mont[i]; a.l[2 * i + 1] = (uint32_t)(mont[i] >> 32); }
            a = cqb::fp_from_mont<cqb::FqP>(a);
            memcpy(out, a.l, 32);
        } else {
            cqb::Fr a;
            for (int i = 0; i < 4; i++) { a.l[2 * i] = (uint32_t)mont[i]; a.l[2 * i + 1] = (uint32_t)(mont[i] >> 32); }
            a = cqb::fp_from_mont<cqb::FrP>(a);
            memcpy(out, a.l, 32);
        }
    }
    void common_scalar(const Fr& s) override {  // transcript.rs:231-236
        uint8_t b[32];
        canonical(s.l, false, b);
        absorbed.push_back(2);
        absorbed.insert(absorbed.end(), b, b + 32);
    }
    void write_scalar(const Fr& s) override {  // :204-208
        common_scalar(s);
        uint8_t b[32];
        canonical(s.l, false, b);
        proof.insert(proof.end(), b, b + 32);
    }
    void write_point(const G1Affine& p) override {  // :199-203, :217-229; compressed encoding derive/curve.rs:635-646
        uint8_t x[32], y[32];
        canonical(p.x, true, x);
        canonical(p.y, true, y);
        absorbed.push_back(1);
        absorbed.insert(absorbed.end(), x, x + 32);
        absorbed.insert(absorbed.end(), y, y + 32);
        x[31] |= (uint8_t)((y[0] & 1) << 7);
        proof.insert(proof.end(), x, x + 32);
    }
    Fr squeeze_challenge_scalar() override {  // :208-215 + Challenge255::new :297-309: from_bytes_wide of the 64-byte digest
        absorbed.push_back(0);
        uint8_t d[64];
        blake2b_512_personal(absorbed, "Halo2-Transcript", d);
        // 512-bit little-endian integer mod r = lo * R^2 * R^-1 ... = mont(lo) + mont(hi) * 2^256: two Montgomery conversions
        cqb::Fr lo, hi;
        memcpy(lo.l, d, 32);
        memcpy(hi.l, d + 32, 32);
        cqb::Fr r = cqb::fp_add<cqb::FrP>(cqb::fp_mul<cqb::FrP>(lo, cqb::Fr::r2()), cqb::fp_mul<cqb::FrP>(hi, cqb::Fr::r3()));
        return detail::from_f(r);  // Montgomery form of (digest mod r), derive/field.rs:29-48 from_u512
    }
};

template <class T>
static std::vector<T> rd(std::ifstream& f, size_t count) {
    std::vector<T> v(count);
    f.read((char*)v.data(), (std::streamsize)(count * sizeof(T)));
    if (!f) { fprintf(stderr, "short circuit file\n"); exit(2); }
    return v;
}

int main(int argc, char** argv) {
    if (argc == 3 && std::string(argv[1]) == "--blake2b") {  // self-check of the hash against hashlib (no GPU): message length in bytes
        std::vector<uint8_t> msg((size_t)atoi(argv[2]));
        for (size_t i = 0; i < msg.size(); i++) msg[i] = (uint8_t)(i * 7 + 3);
        uint8_t d[64];
        blake2b_512_personal(msg, "Halo2-Transcript", d);
        for (int i = 0; i < 64; i++) printf("%02x", d[i]);
        printf("\n");
        return 0;
    }
    if (argc < 3) { fprintf(stderr, "usage: test_create_proof <circuit.bin> <proof.out>\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    auto hdr = rd<uint64_t>(f, 8);  // k, N, A, blinding_factors, cs_degree, m (support size), reserved x2
    const uint32_t k = (uint32_t)hdr[0];
    const size_t N = hdr[1], A = hdr[2], bf = hdr[3], cs_degree = hdr[4], m = hdr[5], n = (size_t)1 << k;
    const size_t chunk = cs_degree - 2, nsets = (A + chunk - 1) / chunk;
    Fr s = rd<Fr>(f, 1)[0], vk_repr = rd<Fr>(f, 1)[0];
    std::vector<std::vector<Fr>> tables_v, advice, sigma, blinds;
    for (int t = 0; t < 2; t++) tables_v.push_back(rd<Fr>(f, N));
    for (size_t j = 0; j < A; j++) advice.push_back(rd<Fr>(f, n));
    for (size_t j = 0; j < A; j++) sigma.push_back(rd<Fr>(f, n));
    for (size_t j = 0; j < nsets; j++) blinds.push_back(rd<Fr>(f, bf));
    std::vector<Fr> rnd = rd<Fr>(f, n);
    plonk::SparseM ms;
    ms.idx = rd<uint32_t>(f, m);
    if (m & 1) rd<uint32_t>(f, 1);  // padding to 8 bytes
    ms.mult = rd<Fr>(f, m);
    init(0);
    {
        ParamsKZG params = ParamsKZG::setup_from_toxic_waste(k, s);
        const size_t Nt = N > n ? N : n;
        TableSRS tsrs = TableSRS::setup_from_toxic_waste(N - 1, s);
        // pk.b0_g1_bound: the last n - 1 powers of the length-Nt table SRS (my_test.rs:205, static_lookup.rs:149)
        std::vector<G1Affine> bound;
        {
            TableSRS big = TableSRS::setup_from_toxic_waste(Nt - 1, s);
            std::vector<G1Affine> g1 = big.download(0);
            bound.assign(g1.begin() + (Nt - (n - 1)), g1.end());
        }
        cqb_bases_t hb = 0;
        detail::check(cqb_bases_register((const uint64_t*)bound.data(), bound.size(), &hb), "register b0 bound");
        StaticTableValues t0(tables_v[0], tsrs), t1(tables_v[1], tsrs);
        plonk::StaticLookup lk{{0, 1}, &tsrs, {&t0, &t1}, hb};
        std::vector<size_t> cols(A);
        std::vector<std::pair<size_t, int>> queries;
        for (size_t j = 0; j < A; j++) { cols[j] = j; queries.push_back({j, 0}); }
        plonk::ProvingKey pk(params, k, cs_degree, bf, cols, sigma, queries, {lk}, vk_repr);
        plonk::ProofRng rng{blinds, rnd};
        Blake2bTranscript t, t2;
        plonk::create_proof(pk, advice, {ms}, rng, t);
        plonk::create_proof(pk, advice, {ms}, rng, t2);  // the pooled working memory is reused: the second proof must be the same
        if (t.proof != t2.proof) { fprintf(stderr, "second proof differs from the first\n"); return 1; }
        std::ofstream o(argv[2], std::ios::binary);
        o.write((const char*)t.proof.data(), (std::streamsize)t.proof.size());
        cqb_bases_free(hb);
    }
    cqb_shutdown();
    printf("ALL OK\n");
    return 0;
}

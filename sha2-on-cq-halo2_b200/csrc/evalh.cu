// evalh.cu — SURVEY.md §8(f) row 1: the row-wise quotient evaluation of plonk::Evaluator::evaluate_h (reference
// halo2_proofs/src/plonk/evaluation.rs:285-551) on the device, so that the coset evaluations produced by
// coeff_to_extended never leave HBM between the coset NTTs (a8) and extended_to_coeff (a9):
//   * graph_eval_kernel   — GraphEvaluator::evaluate (:718-775): one thread per extended-domain row interprets the
//     serialised calculation list (Add/Sub/Mul/Square/Double/Negate/Horner/Store over ValueSources, :41-193). Control flow
//     is uniform across a warp (same program, different row); intermediates live in per-thread local memory, which the
//     hardware interleaves across lanes, so their traffic is coalesced.
//   * cq_lookup_h_kernel  — the static-lookup (CQ) term (:533-548).
//   * permutation_h_kernel — the permutation argument terms (:376-452).
//   * lookup_h_kernel     — the plookup terms (:458-531); table_value comes from the lookup's own GraphEvaluator run.
#include <vector>

#include "internal.h"

namespace cqb {

__device__ __forceinline__ Fr h_ld(const uint4* p, size_t i) {
    uint4 a = __ldg(p + 2 * i), b = __ldg(p + 2 * i + 1);
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ Fr h_ld_rw(const uint4* p, size_t i) {
    uint4 a = p[2 * i], b = p[2 * i + 1];
    Fr r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void h_st(uint4* p, size_t i, const Fr& v) {
    p[2 * i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    p[2 * i + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

constexpr int EVAL_MAX_ROT = 32;

struct GraphArgs {
    const uint4* constants;
    const int* rotations;
    uint32_t n_rot;
    const uint32_t* code;
    uint32_t n_calc;
    const uint4* const* fixed;
    const uint4* const* advice;
    const uint4* const* instance;
    const uint4* challenges;
    Fr beta, gamma, theta, y;
    uint4* values;
    unsigned long long size;
    int rot_scale;
};

template <int MAXI>
__device__ __forceinline__ Fr vs_get(const GraphArgs& a, const uint32_t* w, const uint32_t* rots, const Fr* inter, const Fr& prev) {
    const uint32_t kind = w[0] & 0xffu, rot = w[0] >> 8, idx = w[1];
    switch (kind) {  // ValueSource::get, evaluation.rs:69-109
        case 0: return h_ld(a.constants, idx);
        case 1: return inter[idx];
        case 2: return h_ld(a.fixed[idx], rots[rot]);
        case 3: return h_ld(a.advice[idx], rots[rot]);
        case 4: return h_ld(a.instance[idx], rots[rot]);
        case 5: return h_ld(a.challenges, idx);
        case 6: return a.beta;
        case 7: return a.gamma;
        case 8: return a.theta;
        case 9: return a.y;
        default: return prev;
    }
}

template <int MAXI>
__global__ void __launch_bounds__(128) graph_eval_kernel(const __grid_constant__ GraphArgs a) {
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.size) return;
    uint32_t rots[EVAL_MAX_ROT];
    Fr inter[MAXI];
    for (uint32_t r = 0; r < a.n_rot; r++) {  // get_rotation_idx, evaluation.rs:37-39 (rem_euclid)
        long long v = ((long long)idx + (long long)a.rotations[r] * a.rot_scale) % (long long)a.size;
        if (v < 0) v += (long long)a.size;
        rots[r] = (uint32_t)v;
    }
    const Fr prev = h_ld_rw(a.values, idx);
    Fr last = Fr::zero();
    const uint32_t* pc = a.code;
    for (uint32_t c = 0; c < a.n_calc; c++) {
        const uint32_t op = pc[0], target = pc[1];
        Fr r;
        switch (op) {  // Calculation::evaluate, evaluation.rs:135-193
            case 0: r = fp_add<FrP>(vs_get<MAXI>(a, pc + 2, rots, inter, prev), vs_get<MAXI>(a, pc + 4, rots, inter, prev)); pc += 6; break;
            case 1: r = fp_sub<FrP>(vs_get<MAXI>(a, pc + 2, rots, inter, prev), vs_get<MAXI>(a, pc + 4, rots, inter, prev)); pc += 6; break;
            case 2: r = fp_mul<FrP>(vs_get<MAXI>(a, pc + 2, rots, inter, prev), vs_get<MAXI>(a, pc + 4, rots, inter, prev)); pc += 6; break;
            case 3: { Fr v = vs_get<MAXI>(a, pc + 2, rots, inter, prev); r = fp_mul<FrP>(v, v); pc += 4; break; }
            case 4: r = fp_dbl<FrP>(vs_get<MAXI>(a, pc + 2, rots, inter, prev)); pc += 4; break;
            case 5: r = fp_neg<FrP>(vs_get<MAXI>(a, pc + 2, rots, inter, prev)); pc += 4; break;
            case 6: {  // Horner(start, parts, factor)
                const Fr factor = vs_get<MAXI>(a, pc + 4, rots, inter, prev);
                r = vs_get<MAXI>(a, pc + 2, rots, inter, prev);
                const uint32_t np = pc[6];
                for (uint32_t p = 0; p < np; p++) r = fp_add<FrP>(fp_mul<FrP>(r, factor), vs_get<MAXI>(a, pc + 7 + 2 * p, rots, inter, prev));
                pc += 7 + 2 * np;
                break;
            }
            default: r = vs_get<MAXI>(a, pc + 2, rots, inter, prev); pc += 4; break;  // Store
        }
        inter[target] = r;
        last = r;
    }
    h_st(a.values, idx, last);
}

// evaluation.rs:533-548
__global__ void cq_lookup_h_kernel(uint4* __restrict__ values, const uint4* __restrict__ b, const uint4* __restrict__ f,
                                   const uint4* __restrict__ l_active, Fr beta, Fr y, size_t size) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= size) return;
    Fr v = h_ld_rw(values, i);
    Fr t = fp_sub<FrP>(fp_mul<FrP>(h_ld(b, i), fp_add<FrP>(fp_mul<FrP>(h_ld(f, i), h_ld(l_active, i)), beta)), Fr::one());
    h_st(values, i, fp_add<FrP>(fp_mul<FrP>(v, y), t));
}

// evaluation.rs:458-531: the five plookup constraints of one lookup argument. table_value[idx] is the value of the lookup's
// GraphEvaluator, (compressed input + beta)(compressed table + gamma) (:238-283), computed by graph_evaluate_run.
__global__ void __launch_bounds__(256) lookup_h_kernel(uint4* __restrict__ values, const uint4* __restrict__ table_value,
                                                       const uint4* __restrict__ product, const uint4* __restrict__ pin, const uint4* __restrict__ ptab,
                                                       const uint4* __restrict__ l0, const uint4* __restrict__ l_last, const uint4* __restrict__ l_active,
                                                       Fr beta, Fr gamma, Fr y, unsigned long long size, int rot_scale) {
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= size) return;
    auto rot_idx = [&](int rot) {  // get_rotation_idx, evaluation.rs:37-39
        long long v = ((long long)idx + (long long)rot * rot_scale) % (long long)size;
        if (v < 0) v += (long long)size;
        return (size_t)v;
    };
    const size_t r_next = rot_idx(1), r_prev = rot_idx(-1);
    const Fr z = h_ld(product, idx), a = h_ld(pin, idx), s = h_ld(ptab, idx);
    const Fr l0v = h_ld(l0, idx), lact = h_ld(l_active, idx);
    const Fr a_minus_s = fp_sub<FrP>(a, s);
    Fr v = h_ld_rw(values, idx);
    v = fp_add<FrP>(fp_mul<FrP>(v, y), fp_mul<FrP>(fp_sub<FrP>(Fr::one(), z), l0v));                                   // l_0 (1 - z)
    v = fp_add<FrP>(fp_mul<FrP>(v, y), fp_mul<FrP>(fp_sub<FrP>(fp_sqr<FrP>(z), z), h_ld(l_last, idx)));               // l_last (z^2 - z)
    Fr left = fp_mul<FrP>(fp_mul<FrP>(h_ld(product, r_next), fp_add<FrP>(a, beta)), fp_add<FrP>(s, gamma));
    Fr right = fp_mul<FrP>(z, h_ld(table_value, idx));
    v = fp_add<FrP>(fp_mul<FrP>(v, y), fp_mul<FrP>(fp_sub<FrP>(left, right), lact));                                   // the product rule
    v = fp_add<FrP>(fp_mul<FrP>(v, y), fp_mul<FrP>(a_minus_s, l0v));                                                   // l_0 (a' - s')
    v = fp_add<FrP>(fp_mul<FrP>(v, y), fp_mul<FrP>(fp_mul<FrP>(a_minus_s, fp_sub<FrP>(a, h_ld(pin, r_prev))), lact));  // (a' - s')(a' - a'(w^-1 X))
    h_st(values, idx, v);
}

struct PermArgs {
    uint4* values;
    unsigned long long size;
    int rot_scale, last_rotation;
    uint32_t chunk_len, nsets, ncols;
    const uint4* const* sets;
    const uint4* const* columns;
    const uint4* const* perm_cosets;
    const uint4 *l0, *l_last, *l_active;
    Fr beta, gamma, y, delta_start, delta;
    Fr ew_pw[28];  // extended_omega^(2^b)
};
// evaluation.rs:376-452
__global__ void __launch_bounds__(128) permutation_h_kernel(const __grid_constant__ PermArgs a) {
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.size) return;
    auto rot_idx = [&](int rot) {
        long long v = ((long long)idx + (long long)rot * a.rot_scale) % (long long)a.size;
        if (v < 0) v += (long long)a.size;
        return (size_t)v;
    };
    const size_t r_next = rot_idx(1), r_last = rot_idx(a.last_rotation);
    Fr v = h_ld_rw(a.values, idx);
    const Fr one = Fr::one();
    const Fr l0 = h_ld(a.l0, idx);
    {
        Fr z0 = h_ld(a.sets[0], idx);
        v = fp_add<FrP>(fp_mul<FrP>(v, a.y), fp_mul<FrP>(fp_sub<FrP>(one, z0), l0));  // :397-399
        Fr zl = h_ld(a.sets[a.nsets - 1], idx);
        v = fp_add<FrP>(fp_mul<FrP>(v, a.y), fp_mul<FrP>(fp_sub<FrP>(fp_mul<FrP>(zl, zl), zl), h_ld(a.l_last, idx)));  // :402-406
    }
    for (uint32_t s = 1; s < a.nsets; s++)  // :409-417
        v = fp_add<FrP>(fp_mul<FrP>(v, a.y), fp_mul<FrP>(fp_sub<FrP>(h_ld(a.sets[s], idx), h_ld(a.sets[s - 1], r_last)), l0));
    // beta_term = extended_omega^idx (:390, :450)
    Fr beta_term = Fr::one();
    for (int b = 0; b < 28; b++)
        if ((idx >> b) & 1ull) beta_term = fp_mul<FrP>(beta_term, a.ew_pw[b]);
    Fr current_delta = fp_mul<FrP>(a.delta_start, beta_term);  // :423
    const Fr l_act = h_ld(a.l_active, idx);
    for (uint32_t s = 0; s < a.nsets; s++) {  // :424-449
        const uint32_t c0 = s * a.chunk_len, c1 = min(c0 + a.chunk_len, a.ncols);
        Fr left = h_ld(a.sets[s], r_next);
        for (uint32_t c = c0; c < c1; c++)
            left = fp_mul<FrP>(left, fp_add<FrP>(fp_add<FrP>(h_ld(a.columns[c], idx), fp_mul<FrP>(a.beta, h_ld(a.perm_cosets[c], idx))), a.gamma));
        Fr right = h_ld(a.sets[s], idx);
        for (uint32_t c = c0; c < c1; c++) {
            right = fp_mul<FrP>(right, fp_add<FrP>(fp_add<FrP>(h_ld(a.columns[c], idx), current_delta), a.gamma));
            current_delta = fp_mul<FrP>(current_delta, a.delta);
        }
        v = fp_add<FrP>(fp_mul<FrP>(v, a.y), fp_mul<FrP>(fp_sub<FrP>(left, right), l_act));
    }
    h_st(a.values, idx, v);
}

static Scratch g_eval_meta;
void evalh_release_all() { g_eval_meta.release(); }

// packs host-side metadata (constants, rotations, code, pointer tables, challenges) into one device buffer
struct MetaPacker {
    std::vector<unsigned char> buf;
    size_t add(const void* p, size_t bytes) {
        size_t off = (buf.size() + 31) & ~(size_t)31;
        buf.resize(off + bytes);
        if (bytes) memcpy(buf.data() + off, p, bytes);
        return off;
    }
};

int graph_evaluate_run(const cqb_graph_t* g, const void* const* d_fixed, uint32_t n_fixed, const void* const* d_advice, uint32_t n_advice,
                       const void* const* d_instance, uint32_t n_instance, const uint64_t* challenges, uint32_t n_challenges,
                       const uint64_t* beta, const uint64_t* gamma, const uint64_t* theta, const uint64_t* y, void* d_values, uint64_t size,
                       int32_t rot_scale) {
    if (g->n_rotations > (uint32_t)EVAL_MAX_ROT) return fail(CQB_E_BAD_ARG, "graph has %u rotations (max %d)", g->n_rotations, EVAL_MAX_ROT);
    if (g->num_intermediates > 512) return fail(CQB_E_BAD_ARG, "graph has %u intermediates (max 512)", g->num_intermediates);
    if (size == 0) return 0;
    cudaStream_t st = ctx().stream;
    MetaPacker mp;
    size_t o_const = mp.add(g->constants, (size_t)g->n_constants * 32);
    size_t o_rot = mp.add(g->rotations, (size_t)g->n_rotations * 4);
    size_t o_code = mp.add(g->code, (size_t)g->code_words * 4);
    size_t o_fixed = mp.add(d_fixed, (size_t)n_fixed * 8);
    size_t o_adv = mp.add(d_advice, (size_t)n_advice * 8);
    size_t o_inst = mp.add(d_instance, (size_t)n_instance * 8);
    size_t o_chal = mp.add(challenges, (size_t)n_challenges * 32);
    CQB_TRY(g_eval_meta.ensure(mp.buf.size() + 64));
    CQB_CUDA(cudaMemcpyAsync(g_eval_meta.p, mp.buf.data(), mp.buf.size(), cudaMemcpyHostToDevice, st));
    CQB_CUDA(cudaStreamSynchronize(st));  // mp.buf is a host temporary
    char* base = (char*)g_eval_meta.p;
    GraphArgs a;
    a.constants = (const uint4*)(base + o_const);
    a.rotations = (const int*)(base + o_rot);
    a.n_rot = g->n_rotations;
    a.code = (const uint32_t*)(base + o_code);
    a.n_calc = g->n_calculations;
    a.fixed = (const uint4* const*)(base + o_fixed);
    a.advice = (const uint4* const*)(base + o_adv);
    a.instance = (const uint4* const*)(base + o_inst);
    a.challenges = (const uint4*)(base + o_chal);
    a.beta = fr_from_u64x4(beta); a.gamma = fr_from_u64x4(gamma); a.theta = fr_from_u64x4(theta); a.y = fr_from_u64x4(y);
    a.values = (uint4*)d_values;
    a.size = size;
    a.rot_scale = rot_scale;
    unsigned grid = (unsigned)((size + 127) / 128);
    if (g->num_intermediates <= 32) graph_eval_kernel<32><<<grid, 128, 0, st>>>(a);
    else if (g->num_intermediates <= 128) graph_eval_kernel<128><<<grid, 128, 0, st>>>(a);
    else graph_eval_kernel<512><<<grid, 128, 0, st>>>(a);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

int cq_lookup_h_run(void* d_values, const void* d_b, const void* d_f, const void* d_l_active, const uint64_t* beta, const uint64_t* y, uint64_t size) {
    if (size == 0) return 0;
    cq_lookup_h_kernel<<<(unsigned)((size + 255) / 256), 256, 0, ctx().stream>>>((uint4*)d_values, (const uint4*)d_b, (const uint4*)d_f,
                                                                                  (const uint4*)d_l_active, fr_from_u64x4(beta), fr_from_u64x4(y), size);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

int lookup_h_run(void* d_values, const void* d_table_value, const void* d_product, const void* d_permuted_input, const void* d_permuted_table,
                 const void* d_l0, const void* d_l_last, const void* d_l_active, const uint64_t* beta, const uint64_t* gamma, const uint64_t* y,
                 uint64_t size, int32_t rot_scale) {
    if (size == 0) return 0;
    lookup_h_kernel<<<(unsigned)((size + 255) / 256), 256, 0, ctx().stream>>>(
        (uint4*)d_values, (const uint4*)d_table_value, (const uint4*)d_product, (const uint4*)d_permuted_input, (const uint4*)d_permuted_table,
        (const uint4*)d_l0, (const uint4*)d_l_last, (const uint4*)d_l_active, fr_from_u64x4(beta), fr_from_u64x4(gamma), fr_from_u64x4(y), size,
        rot_scale);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

int permutation_h_run(void* d_values, uint64_t size, int32_t rot_scale, int32_t last_rotation, uint32_t chunk_len, const void* const* d_sets,
                      uint32_t nsets, const void* const* d_columns, const void* const* d_perm_cosets, uint32_t ncols, const void* d_l0,
                      const void* d_l_last, const void* d_l_active, const uint64_t* beta, const uint64_t* gamma, const uint64_t* y,
                      const uint64_t* extended_omega) {
    if (size == 0 || nsets == 0) return 0;  // evaluation.rs:377 `if !sets.is_empty()`
    cudaStream_t st = ctx().stream;
    MetaPacker mp;
    size_t o_sets = mp.add(d_sets, (size_t)nsets * 8), o_cols = mp.add(d_columns, (size_t)ncols * 8), o_perm = mp.add(d_perm_cosets, (size_t)ncols * 8);
    CQB_TRY(g_eval_meta.ensure(mp.buf.size() + 64));
    CQB_CUDA(cudaMemcpyAsync(g_eval_meta.p, mp.buf.data(), mp.buf.size(), cudaMemcpyHostToDevice, st));
    CQB_CUDA(cudaStreamSynchronize(st));
    char* base = (char*)g_eval_meta.p;
    PermArgs a;
    a.values = (uint4*)d_values;
    a.size = size;
    a.rot_scale = rot_scale;
    a.last_rotation = last_rotation;
    a.chunk_len = chunk_len;
    a.nsets = nsets;
    a.ncols = ncols;
    a.sets = (const uint4* const*)(base + o_sets);
    a.columns = (const uint4* const*)(base + o_cols);
    a.perm_cosets = (const uint4* const*)(base + o_perm);
    a.l0 = (const uint4*)d_l0; a.l_last = (const uint4*)d_l_last; a.l_active = (const uint4*)d_l_active;
    a.beta = fr_from_u64x4(beta); a.gamma = fr_from_u64x4(gamma); a.y = fr_from_u64x4(y);
    const uint64_t zeta_raw[4] = {0xb8ca0b2d36636f23ULL, 0xcc37a73fec2bc5e9ULL, 0x048b6e193fd84104ULL, 0x30644e72e131a029ULL};   // fr.rs:112-117
    const uint64_t delta_raw[4] = {0x870e56bbe533e9a2ULL, 0x5b5f898e5e963f25ULL, 0x64ec26aad4c86e71ULL, 0x09226b6e22c6f0caULL};  // fr.rs:104-109
    a.delta_start = fp_mul<FrP>(a.beta, fp_to_mont<FrP>(fr_from_u64x4(zeta_raw)));  // beta * ZETA (:383)
    a.delta = fp_to_mont<FrP>(fr_from_u64x4(delta_raw));
    Fr w = fr_from_u64x4(extended_omega);
    for (int b = 0; b < 28; b++) { a.ew_pw[b] = w; w = fp_sqr<FrP>(w); }
    permutation_h_kernel<<<(unsigned)((size + 127) / 128), 128, 0, st>>>(a);
    CQB_LAUNCHED();
    CQB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace cqb

// capi.cu — the C ABI of libcqb200.so (include/cqb200.h). Each entry point names the reference interface it replaces
// in the header. No CPU fallback anywhere: without a device every compute call fails with CQB_E_NO_DEVICE.
#include <stdarg.h>

#include <algorithm>
#include <map>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "internal.h"

namespace cqb {

static Ctx g_ctxs[MAX_DEVICES];
static int g_nslots = 0;                 // initialised slots: 1 after cqb_init, n after cqb_init_multi(n)
static thread_local int tl_slot = 0;
int cur_slot() { return tl_slot; }
void bind_slot(int slot) {
    tl_slot = slot;
    if (g_ctxs[slot].inited) cudaSetDevice(g_ctxs[slot].device);
}
Ctx& ctx() { return g_ctxs[tl_slot]; }
#define g_ctx (ctx())
static std::recursive_mutex g_mu;
static thread_local char g_errbuf[512];

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_errbuf, sizeof(g_errbuf), fmt, ap);
    va_end(ap);
    g_ctx.last_error = g_errbuf;
    return code;
}

int Scratch::ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) {
        cudaStreamSynchronize(g_ctx.stream);
        cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    size_t want = bytes + bytes / 8;  // headroom so slightly larger follow-up calls do not reallocate
    if (cudaMalloc(&p, want) != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            p = nullptr;
            return fail(CQB_E_OOM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        }
    }
    cap = want;
    return 0;
}
void Scratch::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
int Pinned::ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        p = nullptr;
        return fail(CQB_E_OOM, "cudaMallocHost(%zu) failed", bytes);
    }
    cap = bytes;
    return 0;
}
void Pinned::release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
}

// a registered base set lives on one slot; a SHARDED set (cqb_bases_register_sharded) is a parent whose children hold
// contiguous point ranges on the slots of cqb_init_multi
struct BaseSet {
    void* d; size_t n; bool owned; void* table; int table_c;
    int slot = 0;
    std::vector<cqb_bases_t> shards;   // children (empty for a plain set)
    std::vector<size_t> shard_start;   // first point of each child
};
static std::map<cqb_bases_t, BaseSet> g_bases;
// resident copies of host base slices seen by cqb_msm_bn254_g1_host (see there)
struct HostBasesEntry { const void* ptr; size_t n; uint64_t fp; cqb_bases_t h; unsigned uses; unsigned long long last; size_t bytes; };
static std::vector<HostBasesEntry> g_hb_cache;
static unsigned long long g_hb_clock = 0;
static long long g_hb_budget = -1;  // -1: not decided yet
static cqb_bases_t g_next_handle = 1;
static PerDevice<Scratch> g_scalars, g_idx, g_tmp_bases, g_out;
static Scratch g_io;
static PerDevice<Pinned> g_out_host;
struct CopyStream { cudaStream_t s = nullptr; cudaEvent_t ev[8]; };  // H2D of part p+1 overlaps the kernels of part p (host-pointer MSM)
static PerDevice<CopyStream> g_copy;
static PerDevice<Pinned> g_stage;  // pinned staging ring for pageable host scalars

static int require_init() {
    if (!g_ctx.inited) return fail(CQB_E_NO_DEVICE, "cqb_init() has not been called or no CUDA device is available (there is no CPU fallback)");
    return 0;
}

static int fetch_result(uint64_t out_xy[8], int* is_inf) {
    CQB_TRY(g_out_host->ensure(128));
    CQB_CUDA(cudaMemcpyAsync(g_out_host->p, g_out->p, 80, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    memcpy(out_xy, g_out_host->p, 64);
    if (is_inf) *is_inf = (int)((uint32_t*)g_out_host->p)[16];
    return 0;
}

static Fr fr_arg(const uint64_t* p) { return fr_from_u64x4(p); }

}  // namespace cqb

using namespace cqb;
#define LOCK std::lock_guard<std::recursive_mutex> _lk(g_mu)

extern "C" {

int cqb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// bring up one slot on one device (the caller holds the lock and has bound the slot)
static int init_slot(int slot, int device) {
    tl_slot = slot;
    Ctx& c = g_ctxs[slot];
    CQB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CQB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(CQB_E_NO_DEVICE, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    c.device = device;
    c.sm_count = prop.multiProcessorCount;
    if (const char* v = getenv("CQB_L2_FETCH")) {  // experiment: DRAM -> L2 fetch granularity (32 / 64 / 128 B)
        size_t before = 0, after = 0;
        cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(v));
        cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "cqb: L2 fetch granularity %zu -> %zu (%s)\n", before, after, cudaGetErrorString(e));
        cudaGetLastError();
    }
    CQB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.own_stream = true;
    c.inited = true;
    c.launches = 0;
    CQB_TRY(g_out->ensure(256));
    return 0;
}

int cqb_init(int device) {
    LOCK;
    tl_slot = 0;
    if (g_nslots == 1 && g_ctxs[0].inited && g_ctxs[0].device == device) return 0;
    int n = cqb_device_count();
    if (n <= 0) return fail(CQB_E_NO_DEVICE, "no CUDA device visible (there is no CPU fallback)");
    if (device < 0 || device >= n) return fail(CQB_E_BAD_ARG, "device %d out of range (%d visible)", device, n);
    if (g_nslots) cqb_shutdown();
    CQB_TRY(init_slot(0, device));
    g_nslots = 1;
    return 0;
}

// SURVEY.md section 8(b) `cqb_init(int n_devices)`: ONE process drives devices 0 .. n_devices-1 (slot i = device i). Slot 0 is
// the primary device: every single-device entry point (NTT, polynomial helpers, plain base sets) runs there; base sets
// registered with cqb_bases_register_sharded are split by point range over all slots and their MSMs run on all of them.
int cqb_init_multi(int n_devices) {
    LOCK;
    tl_slot = 0;
    int n = cqb_device_count();
    if (n <= 0) return fail(CQB_E_NO_DEVICE, "no CUDA device visible (there is no CPU fallback)");
    if (n_devices < 1 || n_devices > n || n_devices > MAX_DEVICES) return fail(CQB_E_BAD_ARG, "cqb_init_multi(%d): %d devices visible, at most %d supported", n_devices, n, MAX_DEVICES);
    if (g_nslots == n_devices && g_ctxs[0].inited && g_ctxs[0].device == 0) return 0;
    if (g_nslots) cqb_shutdown();
    for (int i = 0; i < n_devices; i++) {
        int rc = init_slot(i, i);
        if (rc) { std::string msg = g_ctxs[i].last_error; g_nslots = i; cqb_shutdown(); tl_slot = 0; g_ctxs[0].last_error = msg; return rc; }
    }
    g_nslots = n_devices;
    // peer access from the primary device to the others (the partial results travel with cudaMemcpyPeerAsync either way)
    bind_slot(0);
    for (int i = 1; i < n_devices; i++) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, 0, i) == cudaSuccess && can) { if (cudaDeviceEnablePeerAccess(i, 0) != cudaSuccess) cudaGetLastError(); }
        else cudaGetLastError();
    }
    return 0;
}

int cqb_active_devices(void) { return g_nslots; }

void cqb_shutdown(void) {
    LOCK;
    if (!g_nslots && !g_ctxs[0].inited) return;
    for (int slot = MAX_DEVICES - 1; slot >= 0; slot--) {
        Ctx& c = g_ctxs[slot];
        if (!c.inited) continue;
        tl_slot = slot;
        cudaSetDevice(c.device);
        cudaDeviceSynchronize();
        for (auto it = g_bases.begin(); it != g_bases.end();) {
            if (it->second.slot == slot) {
                if (it->second.owned && it->second.d) cudaFree(it->second.d);
                if (it->second.table) cudaFree(it->second.table);
                it = g_bases.erase(it);
            } else ++it;
        }
        g_scalars->release(); g_idx->release(); g_tmp_bases->release(); g_out->release();
        g_out_host->release();
        g_stage->release();
        msm_release_all();
        if (slot == 0) {
            g_io.release();
            ntt_release_all();
            gen_release_all();
            srs_release_all();
            ecntt_release_all();
            poly_release_all();
            products_release_all();
            evalh_release_all();
            g2_release_all();
        }
        if (g_copy->s) {
            for (auto& e : g_copy->ev) cudaEventDestroy(e);
            cudaStreamDestroy(g_copy->s);
            g_copy->s = nullptr;
        }
        if (c.own_stream && c.stream) cudaStreamDestroy(c.stream);
        c.stream = nullptr;
        c.own_stream = false;
        c.inited = false;
    }
    g_bases.clear();
    g_hb_cache.clear();
    g_nslots = 0;
    tl_slot = 0;
}

const char* cqb_last_error(void) { return g_ctx.last_error.c_str(); }

int cqb_set_stream(void* cuda_stream) {
    LOCK;
    CQB_TRY(require_init());
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    if (cuda_stream == nullptr) {
        if (!g_ctx.own_stream) {
            CQB_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
            g_ctx.own_stream = true;
        }
        return 0;
    }
    if (g_ctx.own_stream) cudaStreamDestroy(g_ctx.stream);
    g_ctx.stream = (cudaStream_t)cuda_stream;
    g_ctx.own_stream = false;
    return 0;
}

int cqb_sync(void) {
    LOCK;
    CQB_TRY(require_init());
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return 0;
}

unsigned long long cqb_launch_count(void) {
    unsigned long long t = 0;
    for (int i = 0; i < MAX_DEVICES; i++) t += g_ctxs[i].launches;
    return t;
}

// ---- bases ------------------------------------------------------------------------------------------------------
int cqb_bases_register(const uint64_t* affine_xy, size_t n, cqb_bases_t* out) {
    LOCK;
    CQB_TRY(require_init());
    if (!out || (!affine_xy && n)) return fail(CQB_E_BAD_ARG, "cqb_bases_register: NULL argument");
    void* d = nullptr;
    if (cudaMalloc(&d, n ? n * 64 : 64) != cudaSuccess) {
        cudaGetLastError();
        msm_release_scratch();
        if (cudaMalloc(&d, n ? n * 64 : 64) != cudaSuccess) { cudaGetLastError(); return fail(CQB_E_OOM, "cudaMalloc(%zu) for bases failed", n * 64); }
    }
    if (n) CQB_CUDA(cudaMemcpyAsync(d, affine_xy, n * 64, cudaMemcpyHostToDevice, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    cqb_bases_t h = g_next_handle++;
    g_bases[h] = BaseSet{d, n, true, nullptr, 0};
    *out = h;
    return 0;
}
int cqb_bases_register_device(const void* d_affine_xy, size_t n, cqb_bases_t* out) {
    LOCK;
    CQB_TRY(require_init());
    if (!out || !d_affine_xy) return fail(CQB_E_BAD_ARG, "cqb_bases_register_device: NULL argument");
    cqb_bases_t h = g_next_handle++;
    g_bases[h] = BaseSet{const_cast<void*>(d_affine_xy), n, false, nullptr, 0};
    *out = h;
    return 0;
}
// Split n points by contiguous range over the slots of cqb_init_multi (balanced: the first n % slots shards hold one point more),
// upload every shard to its device; `out` addresses the whole set: MSMs over it run on all devices (cqb_msm_bn254_g1,
// cqb_msm_bn254_g1_multi_dev). With one slot this is cqb_bases_register.
int cqb_bases_register_sharded(const uint64_t* affine_xy, size_t n, cqb_bases_t* out) {
    LOCK;
    CQB_TRY(require_init());
    if (!out || (!affine_xy && n)) return fail(CQB_E_BAD_ARG, "cqb_bases_register_sharded: NULL argument");
    if (g_nslots <= 1) return cqb_bases_register(affine_xy, n, out);
    BaseSet parent{nullptr, n, false, nullptr, 0};
    const size_t base = n / g_nslots, rem = n % g_nslots;
    size_t start = 0;
    int rc = 0;
    for (int i = 0; i < g_nslots && rc == 0; i++) {
        const size_t cnt = base + ((size_t)i < rem ? 1 : 0);
        bind_slot(i);
        void* d = nullptr;
        if (cudaMalloc(&d, cnt ? cnt * 64 : 64) != cudaSuccess) { cudaGetLastError(); rc = fail(CQB_E_OOM, "cudaMalloc(%zu) for a bases shard on device %d failed", cnt * 64, ctx().device); break; }
        cudaError_t e = cnt ? cudaMemcpyAsync(d, affine_xy + start * 8, cnt * 64, cudaMemcpyHostToDevice, g_ctx.stream) : cudaSuccess;
        if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx.stream);
        if (e != cudaSuccess) { cudaFree(d); rc = fail(CQB_E_CUDA, "upload of a bases shard to device %d failed: %s", ctx().device, cudaGetErrorString(e)); break; }
        cqb_bases_t h = g_next_handle++;
        BaseSet child{d, cnt, true, nullptr, 0};
        child.slot = i;
        g_bases[h] = child;
        parent.shards.push_back(h);
        parent.shard_start.push_back(start);
        start += cnt;
    }
    std::string msg = ctx().last_error;
    bind_slot(0);
    if (rc) {
        for (cqb_bases_t h : parent.shards) cqb_bases_free(h);
        g_ctxs[0].last_error = msg;
        return rc;
    }
    cqb_bases_t h = g_next_handle++;
    g_bases[h] = parent;
    *out = h;
    return 0;
}
int cqb_bases_free(cqb_bases_t h) {
    LOCK;
    auto it = g_bases.find(h);
    if (it == g_bases.end()) return fail(CQB_E_BAD_ARG, "unknown bases handle %llu", (unsigned long long)h);
    if (!it->second.shards.empty()) {
        std::vector<cqb_bases_t> kids = it->second.shards;
        g_bases.erase(it);
        for (cqb_bases_t k : kids) cqb_bases_free(k);
        return 0;
    }
    bind_slot(it->second.slot);
    cudaStreamSynchronize(g_ctx.stream);
    if (it->second.owned) cudaFree(it->second.d);
    if (it->second.table) cudaFree(it->second.table);
    g_bases.erase(it);
    bind_slot(0);
    return 0;
}
int cqb_bases_download(cqb_bases_t h, size_t offset, size_t n, uint64_t* affine_xy_out) {
    LOCK;
    CQB_TRY(require_init());
    auto it = g_bases.find(h);
    if (it == g_bases.end()) return fail(CQB_E_BAD_ARG, "unknown bases handle %llu", (unsigned long long)h);
    if (!affine_xy_out && n) return fail(CQB_E_BAD_ARG, "cqb_bases_download: NULL argument");
    if (!it->second.shards.empty()) {
        const BaseSet& ps = it->second;
        if (offset > ps.n || n > ps.n - offset) return fail(CQB_E_LEN_MISMATCH, "download of %zu points at offset %zu exceeds the %zu registered bases", n, offset, ps.n);
        for (size_t i = 0; i < ps.shards.size(); i++) {
            const BaseSet& ch = g_bases[ps.shards[i]];
            const size_t lo = std::max(offset, ps.shard_start[i]), hi = std::min(offset + n, ps.shard_start[i] + ch.n);
            if (hi <= lo) continue;
            bind_slot(ch.slot);
            cudaError_t e = cudaMemcpy(affine_xy_out + (lo - offset) * 8, (const char*)ch.d + (lo - ps.shard_start[i]) * 64, (hi - lo) * 64, cudaMemcpyDeviceToHost);
            bind_slot(0);
            if (e != cudaSuccess) return fail(CQB_E_CUDA, "download from a bases shard failed: %s", cudaGetErrorString(e));
        }
        return 0;
    }
    if (offset > it->second.n || n > it->second.n - offset) return fail(CQB_E_LEN_MISMATCH, "download of %zu points at offset %zu exceeds the %zu registered bases", n, offset, it->second.n);
    if (n) CQB_CUDA(cudaMemcpyAsync(affine_xy_out, (const char*)it->second.d + offset * 64, n * 64, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return 0;
}
int cqb_bases_copy_dev(cqb_bases_t h, size_t offset, size_t n, void* d_affine_xy_out) {
    LOCK;
    CQB_TRY(require_init());
    auto it = g_bases.find(h);
    if (it == g_bases.end()) return fail(CQB_E_BAD_ARG, "unknown bases handle %llu", (unsigned long long)h);
    if (!d_affine_xy_out && n) return fail(CQB_E_BAD_ARG, "cqb_bases_copy_dev: NULL argument");
    if (!it->second.shards.empty()) return fail(CQB_E_BAD_ARG, "cqb_bases_copy_dev: not available for a sharded base set (use cqb_bases_download)");
    if (offset > it->second.n || n > it->second.n - offset) return fail(CQB_E_LEN_MISMATCH, "copy of %zu points at offset %zu exceeds the %zu registered bases", n, offset, it->second.n);
    if (n) CQB_CUDA(cudaMemcpyAsync(d_affine_xy_out, (const char*)it->second.d + offset * 64, n * 64, cudaMemcpyDeviceToDevice, g_ctx.stream));
    return 0;
}
size_t cqb_bases_len(cqb_bases_t h) {
    LOCK;
    auto it = g_bases.find(h);
    return it == g_bases.end() ? 0 : it->second.n;
}

// ---- MSM --------------------------------------------------------------------------------------------------------
static int find_bases(cqb_bases_t h, size_t offset, size_t n, BaseSet** out) {
    auto it = g_bases.find(h);
    if (it == g_bases.end()) return fail(CQB_E_BAD_ARG, "unknown bases handle %llu", (unsigned long long)h);
    // commit / commit_lagrange: assert!(self.n() >= size) poly/kzg/commitment.rs:502,541
    if (offset > it->second.n || n > it->second.n - offset)
        return fail(CQB_E_LEN_MISMATCH, "MSM of %zu scalars at offset %zu exceeds the %zu registered bases", n, offset, it->second.n);
    *out = &it->second;
    return 0;
}

// picks the layout: the precomputed single-set table whenever the set has one, else the windowed layout on the plain bases.
// (Round 1 kept short MSMs — below 1/8 of the set — on the windowed layout because the table's bucket count is sized for the whole
// set; but a windowed MSM always ends in ~254 dependent doublings for the window combination, 1.8 ms on one thread, while the table
// layout has none and its bucket reduction costs 0.2 ms at 2^13 buckets, 0.65 ms at 2^19: the table wins at every length.)
static int dispatch_msm(BaseSet* bs, size_t offset, const void* d_scalars, const uint32_t* d_idx, size_t n, void* d_out = nullptr) {
    if (!d_out) d_out = g_out->p;
    if (bs->table && n)
        return msm_run_precomputed(bs->table, bs->n, bs->table_c, offset, d_scalars, d_idx, n, d_out);
    return msm_run(bs->d, offset, d_scalars, d_idx, n, d_out);
}

int cqb_bases_precompute(cqb_bases_t h, int window_bits) {
    LOCK;
    CQB_TRY(require_init());
    auto it = g_bases.find(h);
    if (it == g_bases.end()) return fail(CQB_E_BAD_ARG, "unknown bases handle %llu", (unsigned long long)h);
    BaseSet& bs = it->second;
    if (bs.n == 0) return 0;
    if (!bs.shards.empty()) {  // every shard builds the table of its own point range, on its own device
        int c = window_bits ? window_bits : msm_precompute_window_bits((bs.n + bs.shards.size() - 1) / bs.shards.size());
        for (cqb_bases_t k : bs.shards) {
            bind_slot(g_bases[k].slot);
            int rc = cqb_bases_precompute(k, c);
            if (rc) { std::string msg = ctx().last_error; bind_slot(0); g_ctxs[0].last_error = msg; return rc; }
        }
        bind_slot(0);
        return 0;
    }
    int c = window_bits ? window_bits : msm_precompute_window_bits(bs.n);
    if (c < 8 || c > 23) return fail(CQB_E_BAD_ARG, "precompute window bits must be 8..23 (got %d)", c);
    if (bs.table && bs.table_c == c) return 0;
    int nwin = msm_windows_for(c);
    if ((size_t)nwin * bs.n >= ((size_t)1 << 31)) return fail(CQB_E_BAD_SIZE, "base set too large for a precomputed table (%zu x %d rows)", bs.n, nwin);
    if (bs.table) { cudaStreamSynchronize(g_ctx.stream); cudaFree(bs.table); bs.table = nullptr; }
    size_t bytes = (size_t)nwin * bs.n * 64;
    size_t free_b = 0, total_b = 0;
    CQB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // Memory plan: the table may take what is free minus the working set an MSM over this set needs afterwards —
    // the bucket-sorted list (nwin x n x 4 B), the affine tree's scratch (72 B per list entry of one part, at most ~16 GiB), the scalars
    // (n x 32 B), histograms / bucket arrays / chunk partials (< 2 GiB) — and a
    // reserve of 1/16 of the device for the caller's polynomials. With window_bits = 0 (automatic) a set whose table does not fit simply
    // stays on the windowed layout (16 windows instead of 13: ~20 % slower, results identical); an explicit window size fails loudly.
    const size_t need_after = bs.n * 32 + msm_working_set_bytes(bs.n, nwin) + total_b / 16;
    if (bytes + need_after > free_b) {  // the grow-only working buffers of earlier, larger calls go first
        msm_release_scratch();
        CQB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    }
    if (bytes + need_after > free_b) {
        if (window_bits == 0) return 0;
        return fail(CQB_E_OOM, "precomputed table needs %zu bytes (+ %zu of MSM working set), only %zu free", bytes, need_after, free_b);
    }
    if (cudaMalloc(&bs.table, bytes) != cudaSuccess) { cudaGetLastError(); bs.table = nullptr; return fail(CQB_E_OOM, "cudaMalloc(%zu) for the precomputed table failed", bytes); }
    bs.table_c = c;
    int rc = msm_precompute_table(bs.d, bs.n, c, bs.table);
    if (rc) { cudaFree(bs.table); bs.table = nullptr; return rc; }
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return 0;
}

int cqb_bases_precomputed_window_bits(cqb_bases_t h) {
    LOCK;
    auto it = g_bases.find(h);
    if (it != g_bases.end() && !it->second.shards.empty()) return cqb_bases_precomputed_window_bits(it->second.shards[0]);
    return (it == g_bases.end() || !it->second.table) ? 0 : it->second.table_c;
}
int cqb_bases_drop_precomputed(cqb_bases_t h) {
    LOCK;
    auto it = g_bases.find(h);
    if (it == g_bases.end()) return fail(CQB_E_BAD_ARG, "unknown bases handle %llu", (unsigned long long)h);
    if (it->second.table) { cudaStreamSynchronize(g_ctx.stream); cudaFree(it->second.table); it->second.table = nullptr; }
    return 0;
}

// ---- host-pointer scalars -> device, pipelined with the MSM ----------------------------------------------------------
// A large MSM from host memory is cut into point-range parts; the transfer of part p+1 overlaps the kernels of part p.
//   pinned memory   : one cudaMemcpyAsync per part on the copy stream.
//   pageable memory : (what a Rust Vec<Fr> is) the driver would stage such a copy synchronously at a few GB/s, so COPY_THREADS
//                     host threads first move the part into a pinned staging buffer (grow-only, one per device slot) and the
//                     asynchronous copy starts from there; the staging of part p+1 runs while the GPU works on part p.
constexpr int COPY_THREADS = 16;
struct HostFeeder : MsmFeeder {
    const uint64_t* src = nullptr;
    bool pinned = true;
    // pageable source: a coordinator thread stages part after part at full memcpy speed from the moment the call starts,
    // independently of how far the enqueuing thread has got; feed() only waits for the part it is about to queue
    std::thread coordinator;
    std::mutex mu;
    std::condition_variable cv;
    int staged = 0;   // parts whose copy has been issued and whose event has been recorded
    int rc = 0;
    std::string err;
    cudaEvent_t* evs = nullptr;

    void start_staging(size_t n, int parts) {
        const int slot = cur_slot();
        char* stage0 = (char*)g_stage->p;
        char* dev0 = (char*)g_scalars->p;
        cudaStream_t cs = g_copy->s;
        evs = g_copy->ev;
        coordinator = std::thread([=] {
            bind_slot(slot);
            size_t bounds[9];
            msm_part_bounds(n, parts, true, bounds);
            // copier threads: the host's cores are shared by every rank of a one-process-per-GPU job (torchrun sets LOCAL_WORLD_SIZE)
            static const int ranks_on_host = std::max(1, getenv("LOCAL_WORLD_SIZE") ? atoi(getenv("LOCAL_WORLD_SIZE")) : 1);
            const int nthreads = std::max(2, std::min(COPY_THREADS, ((int)std::thread::hardware_concurrency() - 2) / ranks_on_host));
            for (int p = 0; p < parts; p++) {
                const size_t lo = bounds[p], cnt = bounds[p + 1] - lo;
                const char* from = (const char*)(src + lo * 4);
                char* stage = stage0 + lo * 32;
                char* dev = dev0 + lo * 32;
                const size_t bytes = cnt * 32, per = (bytes / nthreads + 4095) & ~(size_t)4095;
                // every copier thread stages its slice in four pieces and queues the H2D of a piece as soon as it is staged: the DMA of
                // the first pieces runs under the staging of the later ones
                std::vector<cudaError_t> errs(nthreads, cudaSuccess);
                auto slice = [&](int t) {
                    if (t) bind_slot(slot);
                    const size_t a = std::min(bytes, (size_t)t * per), b = std::min(bytes, (size_t)(t + 1) * per);
                    const size_t piece = ((b - a) / 4 + 4095) & ~(size_t)4095;
                    for (size_t o = a; o < b && piece; o += piece) {
                        const size_t len = std::min(piece, b - o);
                        memcpy(stage + o, from + o, len);
                        cudaError_t e2 = cudaMemcpyAsync(dev + o, stage + o, len, cudaMemcpyHostToDevice, cs);
                        if (e2 != cudaSuccess) { errs[t] = e2; return; }
                    }
                };
                std::vector<std::thread> th;
                for (int t = 1; t < nthreads && per; t++) th.emplace_back(slice, t);
                if (bytes) { if (per) slice(0); else { memcpy(stage, from, bytes); errs[0] = cudaMemcpyAsync(dev, stage, bytes, cudaMemcpyHostToDevice, cs); } }
                for (auto& t : th) t.join();
                cudaError_t e = cudaSuccess;
                for (cudaError_t e2 : errs) if (e2 != cudaSuccess) e = e2;
                if (e == cudaSuccess) e = cudaEventRecord(evs[p & 7], cs);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (e != cudaSuccess) { rc = CQB_E_CUDA; err = std::string("staged copy of host scalars failed: ") + cudaGetErrorString(e); }
                    staged = p + 1;
                }
                cv.notify_all();
                if (e != cudaSuccess) return;
            }
        });
    }
    int feed(int part, size_t lo, size_t cnt, cudaEvent_t* ready_out) override {
        if (pinned) {
            if (cnt) CQB_CUDA(cudaMemcpyAsync((char*)g_scalars->p + lo * 32, src + lo * 4, cnt * 32, cudaMemcpyHostToDevice, g_copy->s));
            CQB_CUDA(cudaEventRecord(g_copy->ev[part & 7], g_copy->s));
            *ready_out = g_copy->ev[part & 7];
            return 0;
        }
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return staged > part || rc != 0; });
        if (rc) return fail(rc, "%s", err.c_str());
        *ready_out = evs[part & 7];
        return 0;
    }
    ~HostFeeder() override {
        if (coordinator.joinable()) coordinator.join();
    }
};

#ifndef CQB_HOST_PARTS
#define CQB_HOST_PARTS 3
#endif
static int g_pageable_parts = 4;  // cqb_msm_set_parts overrides (<= 8: the feeder cycles through 8 events)

// queues `sum scalars[i] * bases[offset + i]` from HOST scalars on the current slot; the result lands in g_out (80 bytes)
static int msm_host_enqueue(BaseSet* bs, size_t offset, const uint64_t* scalars, size_t n, void* d_out = nullptr) {
    if (!d_out) d_out = g_out->p;
    CQB_TRY(g_scalars->ensure(n * 32 + 32));
    bool pinned = false;
    const bool large = n >= ((size_t)1 << 21);
    if (large) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, scalars) == cudaSuccess) pinned = (attr.type == cudaMemoryTypeHost);
        else cudaGetLastError();
    }
    bool staged = large && !pinned && g_pageable_parts > 1 && g_stage->ensure(n * 32) == 0;
    if (large && (pinned || staged)) {
        if (!g_copy->s) {
            CQB_CUDA(cudaStreamCreateWithFlags(&g_copy->s, cudaStreamNonBlocking));
            for (auto& e : g_copy->ev) CQB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        HostFeeder feeder;
        feeder.src = scalars;
        feeder.pinned = pinned;
        static const int env_parts = getenv("CQB_PAGEABLE_PARTS") ? std::max(2, std::min(8, atoi(getenv("CQB_PAGEABLE_PARTS")))) : 0;  // experiments
        // below 2^22 points a part's fixed cost (~0.5 ms of launches and tails) outweighs what a third part hides of the copy
        const int small = n < ((size_t)1 << 22);
        int parts = pinned ? (small ? 2 : CQB_HOST_PARTS) : (env_parts ? env_parts : (small ? 3 : g_pageable_parts));
        // beyond 2^24 points: more parts, so that the largest one (11/16 of the points split over parts - 2, msm_part_bounds) stays near 2^24
        // and the affine tree's scratch near 16 GiB
        if (n > ((size_t)1 << 24)) parts = std::min(8, std::max(parts, 2 + (int)((n / 16 * 11 + ((size_t)1 << 24) - 1) >> 24)));
        if (!pinned) feeder.start_staging(n, parts);
        const bool use_table = bs->table != nullptr;
        if (use_table) return msm_run_precomputed(bs->table, bs->n, bs->table_c, offset, g_scalars->p, nullptr, n, d_out, 1, parts, nullptr, &feeder);
        return msm_run(bs->d, offset, g_scalars->p, nullptr, n, d_out, parts, nullptr, &feeder);
    }
    if (n) CQB_CUDA(cudaMemcpyAsync(g_scalars->p, scalars, n * 32, cudaMemcpyHostToDevice, g_ctx.stream));
    return dispatch_msm(bs, offset, g_scalars->p, nullptr, n, d_out);
}

// ---- MSM over a SHARDED base set: one host thread per device slot (the reference's decomposition across threads,
// arithmetic.rs:137-153), partial points gathered on the primary device with peer copies and folded there -----------------
static PerDevice<cudaEvent_t> g_part_done;
static int msm_sharded(BaseSet* parent, size_t offset, const uint64_t* h_scalars, const void* const* d_scalars_per_shard, size_t n,
                       uint64_t out_xy[8], int* is_inf) {
    const int ns = (int)parent->shards.size();
    std::vector<int> rc(ns, 0);
    std::vector<std::string> msg(ns);
    std::vector<char> active(ns, 0);
    bind_slot(0);
    CQB_TRY(g_tmp_bases->ensure((size_t)ns * 64 + 64));
    CQB_CUDA(cudaMemsetAsync(g_tmp_bases->p, 0, (size_t)ns * 64, g_ctx.stream));  // shards without points contribute the identity
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    void* gather = g_tmp_bases->p;
    const int dev0 = g_ctxs[0].device;
    std::vector<std::thread> th;
    for (int i = 0; i < ns; i++) {
        BaseSet* child = &g_bases[parent->shards[i]];
        const size_t c_lo = parent->shard_start[i], c_hi = c_lo + child->n;
        const size_t lo = std::max(offset, c_lo), hi = std::min(offset + n, c_hi);
        if (hi <= lo) continue;
        active[i] = 1;
        auto work = [=, &rc, &msg]() {
            bind_slot(child->slot);
            int r;
            if (h_scalars) r = msm_host_enqueue(child, lo - c_lo, h_scalars + (lo - offset) * 4, hi - lo);
            else r = dispatch_msm(child, lo - c_lo, d_scalars_per_shard[i], nullptr, hi - lo);
            if (r == 0) {
                cudaEvent_t& ev = g_part_done.get();
                cudaError_t e = cudaSuccess;
                if (!ev) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
                if (e == cudaSuccess) e = cudaMemcpyPeerAsync((char*)gather + (size_t)i * 64, dev0, g_out->p, ctx().device, 64, ctx().stream);
                if (e == cudaSuccess) e = cudaEventRecord(ev, ctx().stream);
                if (e != cudaSuccess) r = fail(CQB_E_CUDA, "gather of the partial point from device %d failed: %s", ctx().device, cudaGetErrorString(e));
            }
            rc[i] = r;
            if (r) msg[i] = ctx().last_error;
        };
        if (child->slot == 0) continue;  // the primary device's shard is queued by this thread below
        th.emplace_back(work);
    }
    // slot 0's own shard on the calling thread
    for (int i = 0; i < ns; i++) {
        BaseSet* child = &g_bases[parent->shards[i]];
        if (!active[i] || child->slot != 0) continue;
        const size_t c_lo = parent->shard_start[i];
        const size_t lo = std::max(offset, c_lo), hi = std::min(offset + n, c_lo + child->n);
        int r;
        if (h_scalars) r = msm_host_enqueue(child, lo - c_lo, h_scalars + (lo - offset) * 4, hi - lo);
        else r = dispatch_msm(child, lo - c_lo, d_scalars_per_shard[i], nullptr, hi - lo);
        if (r == 0) {
            cudaError_t e = cudaMemcpyAsync((char*)gather + (size_t)i * 64, g_out->p, 64, cudaMemcpyDeviceToDevice, g_ctx.stream);
            if (e != cudaSuccess) r = fail(CQB_E_CUDA, "gather of the primary device's partial failed: %s", cudaGetErrorString(e));
        }
        rc[i] = r;
        if (r) msg[i] = ctx().last_error;
    }
    for (auto& t : th) t.join();
    bind_slot(0);
    for (int i = 0; i < ns; i++)
        if (rc[i]) { g_ctxs[0].last_error = msg[i]; return rc[i]; }
    for (int i = 0; i < ns; i++) {
        BaseSet* child = &g_bases[parent->shards[i]];
        if (active[i] && child->slot != 0) CQB_CUDA(cudaStreamWaitEvent(g_ctx.stream, g_part_done.at(child->slot), 0));
    }
    CQB_TRY(g1_sum_affine_run(gather, ns, g_out->p));
    return fetch_result(out_xy, is_inf);
}

int cqb_msm_bn254_g1_dev(cqb_bases_t b, size_t offset, const void* d_scalars, size_t n, uint64_t out_xy[8], int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || (!d_scalars && n)) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_dev: NULL argument");
    BaseSet* bs = nullptr;
    CQB_TRY(find_bases(b, offset, n, &bs));
    if (!bs->shards.empty()) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_dev: sharded base set; use cqb_msm_bn254_g1_multi_dev (one scalar pointer per device)");
    CQB_TRY(dispatch_msm(bs, offset, d_scalars, nullptr, n));
    return fetch_result(out_xy, is_inf);
}

// The same two MSMs with the result LEFT ON THE DEVICE (64 B affine x||y + 16 B identity flag at d_out_xy_flag), queued on the
// library's stream and not waited for: a multi-process caller all-gathers the partial straight from there (sharded.py).
int cqb_msm_bn254_g1_dev_to(cqb_bases_t b, size_t offset, const void* d_scalars, size_t n, void* d_out_xy_flag) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_out_xy_flag || (!d_scalars && n)) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_dev_to: NULL argument");
    BaseSet* bs = nullptr;
    CQB_TRY(find_bases(b, offset, n, &bs));
    if (!bs->shards.empty()) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_dev_to: sharded base set");
    return dispatch_msm(bs, offset, d_scalars, nullptr, n, d_out_xy_flag);
}
int cqb_msm_bn254_g1_to(cqb_bases_t b, size_t offset, const uint64_t* scalars, size_t n, void* d_out_xy_flag) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_out_xy_flag || (!scalars && n)) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_to: NULL argument");
    BaseSet* bs = nullptr;
    CQB_TRY(find_bases(b, offset, n, &bs));
    if (!bs->shards.empty()) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_to: sharded base set");
    return msm_host_enqueue(bs, offset, scalars, n, d_out_xy_flag);
}

// device-resident scalars of a sharded set: d_scalars[i] holds, on shard i's device, the scalars of that shard's point range
// intersected with [offset, offset + n)
int cqb_msm_bn254_g1_multi_dev(cqb_bases_t b, size_t offset, const void* const* d_scalars, size_t n, uint64_t out_xy[8], int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || (!d_scalars && n)) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_multi_dev: NULL argument");
    BaseSet* bs = nullptr;
    CQB_TRY(find_bases(b, offset, n, &bs));
    if (bs->shards.empty()) {
        CQB_TRY(dispatch_msm(bs, offset, d_scalars[0], nullptr, n));
        return fetch_result(out_xy, is_inf);
    }
    return msm_sharded(bs, offset, nullptr, d_scalars, n, out_xy, is_inf);
}

int cqb_msm_bn254_g1(cqb_bases_t b, size_t offset, const uint64_t* scalars, size_t n, uint64_t out_xy[8], int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || (!scalars && n)) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1: NULL argument");
    BaseSet* bs = nullptr;
    CQB_TRY(find_bases(b, offset, n, &bs));
    if (!bs->shards.empty()) return msm_sharded(bs, offset, scalars, nullptr, n, out_xy, is_inf);
    CQB_TRY(msm_host_enqueue(bs, offset, scalars, n));
    return fetch_result(out_xy, is_inf);
}

// B MSMs over the same base range in one pass (commit_lagrange of all advice columns, the h pieces, ...)
static int msm_batch_common(BaseSet* bs, size_t offset, const void* d_scalars, size_t n, int batch, uint64_t* out_xy, int* is_inf) {
    CQB_TRY(g_out->ensure((size_t)batch * 80 + 80));
    CQB_TRY(g_out_host->ensure((size_t)batch * 80 + 80));
    if (bs->table && n && batch > 1) {
        CQB_TRY(msm_run_precomputed(bs->table, bs->n, bs->table_c, offset, d_scalars, nullptr, n, g_out->p, batch));
    } else {
        for (int b = 0; b < batch; b++) {  // no table for this set: one MSM after the other (same results)
            const char* sc = (const char*)d_scalars + (size_t)b * n * 32;
            if (bs->table && n) CQB_TRY(msm_run_precomputed(bs->table, bs->n, bs->table_c, offset, sc, nullptr, n, (char*)g_out->p + (size_t)b * 80));
            else CQB_TRY(msm_run(bs->d, offset, sc, nullptr, n, (char*)g_out->p + (size_t)b * 80));
        }
    }
    CQB_CUDA(cudaMemcpyAsync(g_out_host->p, g_out->p, (size_t)batch * 80, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    for (int b = 0; b < batch; b++) {
        memcpy(out_xy + 8 * b, (char*)g_out_host->p + (size_t)b * 80, 64);
        if (is_inf) is_inf[b] = (int)((uint32_t*)((char*)g_out_host->p + (size_t)b * 80))[16];
    }
    return 0;
}
int cqb_msm_bn254_g1_batch_dev(cqb_bases_t h, size_t offset, const void* d_scalars, size_t n, int batch, uint64_t* out_xy, int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || (!d_scalars && n) || batch < 1 || batch > 64) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_batch_dev: bad argument (1 <= batch <= 64)");
    BaseSet* bs = nullptr;
    CQB_TRY(find_bases(h, offset, n, &bs));
    return msm_batch_common(bs, offset, d_scalars, n, batch, out_xy, is_inf);
}
int cqb_msm_bn254_g1_batch(cqb_bases_t h, size_t offset, const uint64_t* scalars, size_t n, int batch, uint64_t* out_xy, int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || (!scalars && n) || batch < 1 || batch > 64) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_batch: bad argument (1 <= batch <= 64)");
    BaseSet* bs = nullptr;
    CQB_TRY(find_bases(h, offset, n, &bs));
    CQB_TRY(g_scalars->ensure((size_t)batch * n * 32 + 32));
    if (n) CQB_CUDA(cudaMemcpyAsync(g_scalars->p, scalars, (size_t)batch * n * 32, cudaMemcpyHostToDevice, g_ctx.stream));
    return msm_batch_common(bs, offset, g_scalars->p, n, batch, out_xy, is_inf);
}

// ---- the generic best_multiexp(&[Fr], &[G1Affine]) call (arithmetic.rs:132): bases arrive as a host slice on EVERY call -------------
// A prover calls it in loops over the SAME slice (params.g, params.g_lagrange: `commit` per advice column, per h piece, per opening
// witness), and 64 B per point of upload dwarf the 32 B of scalars (1 GiB per call at 2^24). Large host slices are therefore kept
// resident after their first use, keyed by (pointer, length, fingerprint): the fingerprint hashes 4096 spread points plus the first and
// last 64 of the slice (~0.1 ms) — an SRS slice is immutable for the life of the
// params, and any edit that keeps all of those points is not a case this call path has. From the second use on the slice is served by
// a registered set (no upload), from the third by its precomputed table. CQB_HOST_BASES_CACHE=0 (environment) or
// cqb_set_host_bases_cache(0) turns the cache off; its budget is a byte count (default 1/8 of the device), least recently used out first.
constexpr size_t HB_MIN_POINTS = (size_t)1 << 14;

static uint64_t hb_fingerprint(const uint64_t* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull ^ n;
    auto mix = [&](size_t i) {
        for (int k = 0; k < 8; k++) { h ^= p[i * 8 + k]; h *= 0x100000001b3ull; h ^= h >> 29; }
    };
    if (n <= 8192) { for (size_t i = 0; i < n; i++) mix(i); return h; }
    for (size_t i = 0; i < 64; i++) { mix(i); mix(n - 1 - i); }
    const size_t step = n / 4096;
    for (size_t i = 0; i < 4096; i++) mix(i * step + (i % step));
    return h;
}
static void hb_evict_until(size_t need) {
    size_t used = 0;
    for (auto& e : g_hb_cache) used += e.bytes;
    while (!g_hb_cache.empty() && (long long)(used + need) > g_hb_budget) {
        size_t v = 0;
        for (size_t i = 1; i < g_hb_cache.size(); i++) if (g_hb_cache[i].last < g_hb_cache[v].last) v = i;
        used -= g_hb_cache[v].bytes;
        cqb_bases_free(g_hb_cache[v].h);
        g_hb_cache.erase(g_hb_cache.begin() + v);
    }
}

int cqb_set_host_bases_cache(long long budget_bytes) {
    LOCK;
    g_hb_budget = budget_bytes < 0 ? -1 : budget_bytes;
    if (g_hb_budget >= 0) hb_evict_until(0);
    return 0;
}

int cqb_msm_bn254_g1_host(const uint64_t* affine_xy, const uint64_t* scalars, size_t n, uint64_t out_xy[8], int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || ((!scalars || !affine_xy) && n)) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_host: NULL argument");
    if (g_hb_budget < 0) {
        const char* e = getenv("CQB_HOST_BASES_CACHE");
        size_t free_b = 0, total_b = 0;
        if (e) g_hb_budget = atoll(e);
        else g_hb_budget = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess ? (long long)(total_b / 8) : 0;
    }
    if (n >= HB_MIN_POINTS && g_hb_budget > 0 && (long long)(n * 64) <= g_hb_budget) {
        const uint64_t fp = hb_fingerprint(affine_xy, n);
        HostBasesEntry* hit = nullptr;
        for (auto& e : g_hb_cache) if (e.ptr == affine_xy && e.n == n && e.fp == fp) { hit = &e; break; }
        if (!hit) {
            // a stale entry for the same pointer (freed and reallocated slice) goes first
            for (size_t i = 0; i < g_hb_cache.size();) {
                if (g_hb_cache[i].ptr == affine_xy) { cqb_bases_free(g_hb_cache[i].h); g_hb_cache.erase(g_hb_cache.begin() + i); } else i++;
            }
            hb_evict_until(n * 64);
            cqb_bases_t h = 0;
            if (cqb_bases_register(affine_xy, n, &h) == 0) {
                g_hb_cache.push_back(HostBasesEntry{affine_xy, n, fp, h, 0, 0, n * 64});
                hit = &g_hb_cache.back();
            }
        }
        if (hit) {
            hit->uses++;
            hit->last = ++g_hb_clock;
            if (hit->uses == 3 && n >= ((size_t)1 << 16)) {  // third use of the same slice: worth a table (memory-aware, may decline)
                const size_t before = hit->bytes;
                if (cqb_bases_precompute(hit->h, 0) == 0 && cqb_bases_precomputed_window_bits(hit->h) > 0)
                    hit->bytes = before + (size_t)msm_windows_for(cqb_bases_precomputed_window_bits(hit->h)) * n * 64;
            }
            const cqb_bases_t h = hit->h;
            return cqb_msm_bn254_g1(h, 0, scalars, n, out_xy, is_inf);
        }
    }
    CQB_TRY(g_tmp_bases->ensure(n * 64 + 64));
    CQB_TRY(g_scalars->ensure(n * 32 + 32));
    if (n) {
        CQB_CUDA(cudaMemcpyAsync(g_tmp_bases->p, affine_xy, n * 64, cudaMemcpyHostToDevice, g_ctx.stream));
        CQB_CUDA(cudaMemcpyAsync(g_scalars->p, scalars, n * 32, cudaMemcpyHostToDevice, g_ctx.stream));
    }
    CQB_TRY(msm_run(g_tmp_bases->p, 0, g_scalars->p, nullptr, n, g_out->p));
    return fetch_result(out_xy, is_inf);
}

// MSMKZG::eval (poly/kzg/msm.rs:65-70): the bases are PROJECTIVE (E::G1, Jacobian x, y, z; 96 B each): batch_normalize
// (derive/curve.rs:362-397) on the device, then the MSM. Verifier-side and small; one-shot like cqb_msm_bn254_g1_host.
int cqb_g1_batch_normalize(const uint64_t* jacobian_xyz, size_t n, uint64_t* affine_xy_out) {
    LOCK;
    CQB_TRY(require_init());
    if ((!jacobian_xyz || !affine_xy_out) && n) return fail(CQB_E_BAD_ARG, "cqb_g1_batch_normalize: NULL argument");
    CQB_TRY(g_tmp_bases->ensure(n * 160 + 64));
    char* d_jac = (char*)g_tmp_bases->p + n * 64;
    if (n) CQB_CUDA(cudaMemcpyAsync(d_jac, jacobian_xyz, n * 96, cudaMemcpyHostToDevice, g_ctx.stream));
    CQB_TRY(g1_batch_normalize_run(d_jac, n, g_tmp_bases->p));
    if (n) CQB_CUDA(cudaMemcpyAsync(affine_xy_out, g_tmp_bases->p, n * 64, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return 0;
}
int cqb_msm_bn254_g1_jacobian(const uint64_t* jacobian_xyz, const uint64_t* scalars, size_t n, uint64_t out_xy[8], int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || ((!scalars || !jacobian_xyz) && n)) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_jacobian: NULL argument");
    CQB_TRY(g_tmp_bases->ensure(n * 160 + 64));
    CQB_TRY(g_scalars->ensure(n * 32 + 32));
    char* d_jac = (char*)g_tmp_bases->p + n * 64;
    if (n) {
        CQB_CUDA(cudaMemcpyAsync(d_jac, jacobian_xyz, n * 96, cudaMemcpyHostToDevice, g_ctx.stream));
        CQB_CUDA(cudaMemcpyAsync(g_scalars->p, scalars, n * 32, cudaMemcpyHostToDevice, g_ctx.stream));
    }
    CQB_TRY(g1_batch_normalize_run(d_jac, n, g_tmp_bases->p));
    CQB_TRY(msm_run(g_tmp_bases->p, 0, g_scalars->p, nullptr, n, g_out->p));
    return fetch_result(out_xy, is_inf);
}

int cqb_msm_bn254_g1_sparse(cqb_bases_t b, const uint32_t* idx, const uint64_t* scalars, size_t m, uint64_t out_xy[8], int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || ((!scalars || !idx) && m)) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_sparse: NULL argument");
    auto it = g_bases.find(b);
    if (it == g_bases.end()) return fail(CQB_E_BAD_ARG, "unknown bases handle %llu", (unsigned long long)b);
    for (size_t j = 0; j < m; j++)  // the reference would panic on an out-of-range table index (slice indexing)
        if (idx[j] >= it->second.n) return fail(CQB_E_BAD_ARG, "sparse index %u out of range (%zu bases)", idx[j], it->second.n);
    CQB_TRY(g_scalars->ensure(m * 32 + 32));
    CQB_TRY(g_idx->ensure(m * 4 + 4));
    if (m) {
        CQB_CUDA(cudaMemcpyAsync(g_scalars->p, scalars, m * 32, cudaMemcpyHostToDevice, g_ctx.stream));
        CQB_CUDA(cudaMemcpyAsync(g_idx->p, idx, m * 4, cudaMemcpyHostToDevice, g_ctx.stream));
    }
    CQB_TRY(dispatch_msm(&it->second, 0, g_scalars->p, g_idx->as<uint32_t>(), m));
    return fetch_result(out_xy, is_inf);
}

int cqb_g1_sum_affine(const uint64_t* affine_xy, size_t n, uint64_t out_xy[8], int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || (!affine_xy && n)) return fail(CQB_E_BAD_ARG, "cqb_g1_sum_affine: NULL argument");
    CQB_TRY(g_tmp_bases->ensure(n * 64 + 64));
    if (n) CQB_CUDA(cudaMemcpyAsync(g_tmp_bases->p, affine_xy, n * 64, cudaMemcpyHostToDevice, g_ctx.stream));
    CQB_TRY(g1_sum_affine_run(g_tmp_bases->p, n, g_out->p));
    return fetch_result(out_xy, is_inf);
}

int cqb_g1_sum_affine_dev(const void* d_affine_xy, size_t n, uint64_t out_xy[8], int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || (!d_affine_xy && n)) return fail(CQB_E_BAD_ARG, "cqb_g1_sum_affine_dev: NULL argument");
    CQB_TRY(g1_sum_affine_run(d_affine_xy, n, g_out->p));
    return fetch_result(out_xy, is_inf);
}

// ---- NTT --------------------------------------------------------------------------------------------------------
static int check_log_n(uint32_t log_n) {
    if (log_n > 28) return fail(CQB_E_BAD_SIZE, "log_n = %u exceeds Fr::S = 28 (bn256/fr.rs:72)", log_n);
    return 0;
}

int cqb_ntt_bn254_fr_dev(void* d_a, const uint64_t omega[4], uint32_t log_n) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_a || !omega) return fail(CQB_E_BAD_ARG, "cqb_ntt_bn254_fr_dev: NULL argument");
    CQB_TRY(check_log_n(log_n));
    NttFused f;
    return ntt_run(d_a, d_a, log_n, omega, f);
}

// ---- building blocks of the distributed four-step NTT (one process per GPU; the exchange itself is NCCL, in sharded.py) ----
int cqb_ntt_bn254_fr_batch_dev(void* d_a, const uint64_t omega[4], uint32_t log_n, uint32_t batch) {
    LOCK;
    CQB_TRY(require_init());
    if ((!d_a && batch) || !omega) return fail(CQB_E_BAD_ARG, "cqb_ntt_bn254_fr_batch_dev: NULL argument");
    CQB_TRY(check_log_n(log_n));
    NttFused f;
    return ntt_run(d_a, d_a, log_n, omega, f, batch);
}
int cqb_ntt_bn254_fr_batch_map_dev(const void* d_src, void* d_dst, const uint64_t omega[4], uint32_t log_n, uint32_t batch, int in_seg_log,
                                   int out_transposed, const uint64_t tw_omega[4], uint32_t tw_log_n, size_t tw_row0, uint32_t in_batch_total) {
    LOCK;
    CQB_TRY(require_init());
    if (((!d_src || !d_dst) && batch) || !omega || d_src == d_dst) return fail(CQB_E_BAD_ARG, "cqb_ntt_bn254_fr_batch_map_dev: NULL or aliasing argument");
    CQB_TRY(check_log_n(log_n));
    if (in_seg_log > (int)log_n) return fail(CQB_E_BAD_ARG, "segment longer than the transform");
    NttFused f;
    if (in_seg_log >= 0) {
        f.in_map = 1;
        f.in_s = (unsigned)in_seg_log;
        // in_batch_total > batch: this call transforms a GROUP of the members held by the receive buffer (d_src then points at
        // the group's first member inside the first segment)
        f.in_A = (unsigned long long)(in_batch_total ? in_batch_total : batch) << in_seg_log;
        f.in_B = 1ull << in_seg_log;
    }
    if (out_transposed) {
        f.out_map = 1;
        f.out_s = 0;
        f.out_A = batch;
        f.out_B = 1;
    }
    if (tw_omega) {
        if (tw_log_n == 0 || tw_log_n > 28) return fail(CQB_E_BAD_SIZE, "twiddle log_n = %u out of range", tw_log_n);
        const void* t2 = nullptr;
        CQB_TRY(ntt_get_twiddles(tw_omega, tw_log_n, &t2));
        f.tw2 = t2;
        f.tw2_L = tw_log_n;
        f.tw2_row0 = tw_row0;
    }
    return ntt_run(d_src, d_dst, log_n, omega, f, batch);
}
// the same transform with its last store going into the peers' receive buffers (out_map == 2); d_scratch: batch * 2^log_n
// elements of local scratch for the intermediate passes (never written by a peer)
int cqb_ntt_bn254_fr_batch_p2p_dev(const void* d_src, void* d_scratch, void* const* peer_dst, uint32_t n_peers, uint32_t self_rank,
                                   const uint64_t omega[4], uint32_t log_n, uint32_t batch, int in_seg_log, const uint64_t tw_omega[4],
                                   uint32_t tw_log_n, size_t tw_row0) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_src || !d_scratch || !peer_dst || !omega || d_src == d_scratch) return fail(CQB_E_BAD_ARG, "cqb_ntt_bn254_fr_batch_p2p_dev: NULL or aliasing argument");
    if (n_peers < 1 || n_peers > 8 || (n_peers & (n_peers - 1)) || self_rank >= n_peers) return fail(CQB_E_BAD_ARG, "peers: 1, 2, 4 or 8 (got %u)", n_peers);
    CQB_TRY(check_log_n(log_n));
    uint32_t g = 0;
    while ((1u << g) < n_peers) g++;
    if (log_n < g || in_seg_log > (int)log_n) return fail(CQB_E_BAD_ARG, "transform shorter than the peer count / segment");
    NttFused f;
    if (in_seg_log >= 0) {
        f.in_map = 1;
        f.in_s = (unsigned)in_seg_log;
        f.in_A = (unsigned long long)batch << in_seg_log;
        f.in_B = 1ull << in_seg_log;
    }
    f.out_map = 2;
    f.peer_rows_log = log_n - g;
    f.peer_self_off = ((unsigned long long)self_rank << (log_n - g)) * batch;
    for (uint32_t h = 0; h < n_peers; h++) {
        if (!peer_dst[h]) return fail(CQB_E_BAD_ARG, "NULL peer buffer %u", h);
        f.peer[h] = peer_dst[h];
    }
    if (tw_omega) {
        if (tw_log_n == 0 || tw_log_n > 28) return fail(CQB_E_BAD_SIZE, "twiddle log_n = %u out of range", tw_log_n);
        const void* t2 = nullptr;
        CQB_TRY(ntt_get_twiddles(tw_omega, tw_log_n, &t2));
        f.tw2 = t2;
        f.tw2_L = tw_log_n;
        f.tw2_row0 = tw_row0;
    }
    return ntt_run(d_src, d_scratch, log_n, omega, f, batch);
}
// CUDA IPC plumbing for the peer buffers (one process per GPU): export a cqb_dev_alloc'ed buffer, open a peer's, close it
int cqb_ipc_export(const void* d_ptr, unsigned char handle_out[64]) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_ptr || !handle_out) return fail(CQB_E_BAD_ARG, "cqb_ipc_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CQB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle_out, &h, 64);
    return 0;
}
int cqb_ipc_open(const unsigned char handle[64], void** d_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!handle || !d_out) return fail(CQB_E_BAD_ARG, "cqb_ipc_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CQB_CUDA(cudaIpcOpenMemHandle(d_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int cqb_ipc_close(void* d_ptr) {
    LOCK;
    CQB_TRY(require_init());
    if (d_ptr) CQB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}
int cqb_fr_mul_omega_powers_dev(void* d_a, size_t rows, size_t cols, size_t row0, const uint64_t omega[4], uint32_t log_n) {
    LOCK;
    CQB_TRY(require_init());
    if ((!d_a && rows && cols) || !omega) return fail(CQB_E_BAD_ARG, "cqb_fr_mul_omega_powers_dev: NULL argument");
    return fr_mul_omega_powers_run(d_a, rows, cols, row0, omega, log_n);
}
int cqb_fr_transpose_dev(const void* d_in, void* d_out, size_t rows, size_t cols) {
    LOCK;
    CQB_TRY(require_init());
    if ((!d_in || !d_out || d_in == d_out) && rows && cols) return fail(CQB_E_BAD_ARG, "cqb_fr_transpose_dev: NULL or aliasing argument");
    return fr_transpose_run(d_in, d_out, rows, cols);
}

int cqb_intt_bn254_fr_dev(void* d_a, const uint64_t omega_inv[4], const uint64_t divisor[4], uint32_t log_n) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_a || !omega_inv || !divisor) return fail(CQB_E_BAD_ARG, "cqb_intt_bn254_fr_dev: NULL argument");
    CQB_TRY(check_log_n(log_n));
    NttFused f;
    f.post_mode = 1;
    f.post[0] = fr_arg(divisor);
    return ntt_run(d_a, d_a, log_n, omega_inv, f);
}

int cqb_coset_ntt_bn254_fr_dev(const void* d_coeffs, size_t n, void* d_out, const uint64_t ext_omega[4], uint32_t ext_log_n,
                               const uint64_t g_coset[4], const uint64_t g_coset_inv[4]) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_coeffs || !d_out || !ext_omega || !g_coset || !g_coset_inv) return fail(CQB_E_BAD_ARG, "cqb_coset_ntt_bn254_fr_dev: NULL argument");
    CQB_TRY(check_log_n(ext_log_n));
    if (n > ((size_t)1 << ext_log_n)) return fail(CQB_E_BAD_SIZE, "coset NTT: %zu coefficients do not fit 2^%u", n, ext_log_n);
    NttFused f;
    f.n_in = n;
    if (n == 0) {  // all-zero polynomial
        CQB_CUDA(cudaMemsetAsync(d_out, 0, ((size_t)32) << ext_log_n, g_ctx.stream));
        return 0;
    }
    f.pre_mode = 1;  // distribute_powers_zeta(into_coset = true): [1, g_coset, g_coset_inv][i % 3]  domain.rs:347-363
    f.pre[0] = Fr::one();
    f.pre[1] = fr_arg(g_coset);
    f.pre[2] = fr_arg(g_coset_inv);
    return ntt_run(d_coeffs, d_out, ext_log_n, ext_omega, f);
}

static Scratch g_tev;
int cqb_coset_intt_bn254_fr_dev(void* d_a, uint32_t ext_log_n, const uint64_t ext_omega_inv[4], const uint64_t ext_divisor[4],
                                const uint64_t g_coset[4], const uint64_t g_coset_inv[4], const uint64_t* t_evaluations,
                                uint32_t t_len) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_a || !ext_omega_inv || !ext_divisor || !g_coset || !g_coset_inv) return fail(CQB_E_BAD_ARG, "cqb_coset_intt_bn254_fr_dev: NULL argument");
    CQB_TRY(check_log_n(ext_log_n));
    NttFused f;
    if (t_evaluations) {
        if (t_len == 0 || (t_len & (t_len - 1)) || t_len > ((size_t)1 << ext_log_n))
            return fail(CQB_E_BAD_ARG, "t_evaluations length %u must be a power of two <= 2^%u (domain.rs:100)", t_len, ext_log_n);
        if (t_len <= (uint32_t)NTT_PRE_MAX_PUB) {
            f.pre_mode = 2;
            f.pre_len = (int)t_len;
            for (uint32_t i = 0; i < t_len; i++) f.pre[i] = fr_arg(t_evaluations + 4 * i);
        } else {  // long table: separate element-wise pass (divide_by_vanishing_poly, domain.rs:319-338)
            CQB_TRY(g_tev.ensure((size_t)t_len * 32));
            CQB_CUDA(cudaMemcpyAsync(g_tev.p, t_evaluations, (size_t)t_len * 32, cudaMemcpyHostToDevice, g_ctx.stream));
            CQB_TRY(fr_scale_table(d_a, (size_t)1 << ext_log_n, g_tev.p, t_len));
        }
    }
    // ifft divisor and distribute_powers_zeta(into_coset = false) = [1, g_coset_inv, g_coset][i % 3] fused: the field is
    // exact, so (a * divisor) * cp == a * (divisor * cp)
    Fr dv = fr_arg(ext_divisor);
    f.post_mode = 2;
    f.post[0] = dv;
    f.post[1] = fp_mul<FrP>(dv, fr_arg(g_coset_inv));
    f.post[2] = fp_mul<FrP>(dv, fr_arg(g_coset));
    return ntt_run(d_a, d_a, ext_log_n, ext_omega_inv, f);
}

// host-buffer variants: H2D, device path, D2H
static int with_host_buffer(uint64_t* a, size_t n_in, size_t n_out, int (*body)(void* d, void* user), void* user) {
    size_t cap = (n_in > n_out ? n_in : n_out) * 32;
    CQB_TRY(g_io.ensure(cap + 32));
    if (n_in) CQB_CUDA(cudaMemcpyAsync(g_io.p, a, n_in * 32, cudaMemcpyHostToDevice, g_ctx.stream));
    CQB_TRY(body(g_io.p, user));
    if (n_out) CQB_CUDA(cudaMemcpyAsync(a, g_io.p, n_out * 32, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return 0;
}

int cqb_ntt_bn254_fr(uint64_t* a, const uint64_t omega[4], uint32_t log_n) {
    LOCK;
    CQB_TRY(require_init());
    if (!a || !omega) return fail(CQB_E_BAD_ARG, "cqb_ntt_bn254_fr: NULL argument");
    CQB_TRY(check_log_n(log_n));
    struct U { const uint64_t* w; uint32_t l; } u{omega, log_n};
    size_t n = (size_t)1 << log_n;
    return with_host_buffer(a, n, n, [](void* d, void* up) { U* x = (U*)up; return cqb_ntt_bn254_fr_dev(d, x->w, x->l); }, &u);
}

int cqb_intt_bn254_fr(uint64_t* a, const uint64_t omega_inv[4], const uint64_t divisor[4], uint32_t log_n) {
    LOCK;
    CQB_TRY(require_init());
    if (!a || !omega_inv || !divisor) return fail(CQB_E_BAD_ARG, "cqb_intt_bn254_fr: NULL argument");
    CQB_TRY(check_log_n(log_n));
    struct U { const uint64_t *w, *dv; uint32_t l; } u{omega_inv, divisor, log_n};
    size_t n = (size_t)1 << log_n;
    return with_host_buffer(a, n, n, [](void* d, void* up) { U* x = (U*)up; return cqb_intt_bn254_fr_dev(d, x->w, x->dv, x->l); }, &u);
}

static Scratch g_io2;
int cqb_coset_ntt_bn254_fr(const uint64_t* coeffs, size_t n, uint64_t* out, const uint64_t ext_omega[4], uint32_t ext_log_n,
                           const uint64_t g_coset[4], const uint64_t g_coset_inv[4]) {
    LOCK;
    CQB_TRY(require_init());
    if ((!coeffs && n) || !out || !ext_omega || !g_coset || !g_coset_inv) return fail(CQB_E_BAD_ARG, "cqb_coset_ntt_bn254_fr: NULL argument");
    CQB_TRY(check_log_n(ext_log_n));
    size_t en = (size_t)1 << ext_log_n;
    if (n > en) return fail(CQB_E_BAD_SIZE, "coset NTT: %zu coefficients do not fit 2^%u", n, ext_log_n);
    CQB_TRY(g_io2.ensure(n * 32 + 32));
    CQB_TRY(g_io.ensure(en * 32));
    if (n) CQB_CUDA(cudaMemcpyAsync(g_io2.p, coeffs, n * 32, cudaMemcpyHostToDevice, g_ctx.stream));
    CQB_TRY(cqb_coset_ntt_bn254_fr_dev(g_io2.p, n, g_io.p, ext_omega, ext_log_n, g_coset, g_coset_inv));
    CQB_CUDA(cudaMemcpyAsync(out, g_io.p, en * 32, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return 0;
}

int cqb_coset_intt_bn254_fr(uint64_t* a, uint32_t ext_log_n, const uint64_t ext_omega_inv[4], const uint64_t ext_divisor[4],
                            const uint64_t g_coset[4], const uint64_t g_coset_inv[4], const uint64_t* t_evaluations, uint32_t t_len) {
    LOCK;
    CQB_TRY(require_init());
    if (!a) return fail(CQB_E_BAD_ARG, "cqb_coset_intt_bn254_fr: NULL argument");
    CQB_TRY(check_log_n(ext_log_n));
    struct U { uint32_t l; const uint64_t *w, *dv, *g, *gi, *t; uint32_t tl; } u{ext_log_n, ext_omega_inv, ext_divisor, g_coset, g_coset_inv, t_evaluations, t_len};
    size_t n = (size_t)1 << ext_log_n;
    return with_host_buffer(a, n, n, [](void* d, void* up) {
        U* x = (U*)up;
        return cqb_coset_intt_bn254_fr_dev(d, x->l, x->w, x->dv, x->g, x->gi, x->t, x->tl);
    }, &u);
}

// ---- synthetic inputs ------------------------------------------------------------------------------------------------
int cqb_synth_scalars_dev(uint64_t seed, size_t start, size_t n, void* d_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_out && n) return fail(CQB_E_BAD_ARG, "cqb_synth_scalars_dev: NULL argument");
    return synth_scalars_run(seed, start, n, d_out);
}
int cqb_synth_scalars_dev_on(int slot, uint64_t seed, size_t start, size_t n, void* d_out) {
    LOCK;
    CQB_TRY(require_init());
    if (slot < 0 || slot >= g_nslots) return fail(CQB_E_BAD_ARG, "device slot %d out of range (%d active)", slot, g_nslots);
    if (!d_out && n) return fail(CQB_E_BAD_ARG, "cqb_synth_scalars_dev_on: NULL argument");
    bind_slot(slot);
    int rc = synth_scalars_run(seed, start, n, d_out);  // a plain kernel on the slot's stream, no per-device scratch
    if (rc == 0 && cudaStreamSynchronize(ctx().stream) != cudaSuccess) rc = fail(CQB_E_CUDA, "synth scalars on slot %d failed", slot);
    std::string msg = ctx().last_error;
    bind_slot(0);
    if (rc) g_ctxs[0].last_error = msg;
    return rc;
}
int cqb_synth_bases_dev(uint64_t seed, size_t start, size_t n, void* d_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_out && n) return fail(CQB_E_BAD_ARG, "cqb_synth_bases_dev: NULL argument");
    return synth_bases_run(seed, start, n, d_out);
}

// ---- SRS generation and element-wise helpers ----------------------------------------------
int cqb_srs_setup_dev(uint32_t k, const uint64_t s[4], void* d_g, void* d_g_lagrange) {
    LOCK;
    CQB_TRY(require_init());
    if (!s || !d_g || !d_g_lagrange) return fail(CQB_E_BAD_ARG, "cqb_srs_setup_dev: NULL argument");
    return srs_setup_run(k, s, d_g, d_g_lagrange, nullptr);
}
int cqb_table_srs_setup_dev(uint32_t log_len, const uint64_t s[4], void* d_g1, void* d_g1_lagrange, void* d_opening_at_0) {
    LOCK;
    CQB_TRY(require_init());
    if (!s || !d_g1 || !d_g1_lagrange || !d_opening_at_0) return fail(CQB_E_BAD_ARG, "cqb_table_srs_setup_dev: NULL argument");
    return srs_setup_run(log_len, s, d_g1, d_g1_lagrange, d_opening_at_0);
}
int cqb_g_to_lagrange_dev(const void* d_g, uint32_t k, void* d_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_g || !d_out || d_g == d_out) return fail(CQB_E_BAD_ARG, "cqb_g_to_lagrange_dev: NULL or aliasing arguments");
    return g_to_lagrange_run(d_g, k, d_out);
}
int cqb_cq_table_qs_dev(const void* d_table_coeffs, uint32_t log_n, const void* d_srs_g1, void* d_qs_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_table_coeffs || !d_srs_g1 || !d_qs_out) return fail(CQB_E_BAD_ARG, "cqb_cq_table_qs_dev: NULL argument");
    return cq_table_qs_run(d_table_coeffs, log_n, d_srs_g1, d_qs_out);
}
int cqb_g1_generator_mul_dev(const void* d_scalars, size_t n, void* d_out) {
    LOCK;
    CQB_TRY(require_init());
    if ((!d_scalars || !d_out) && n) return fail(CQB_E_BAD_ARG, "cqb_g1_generator_mul_dev: NULL argument");
    return g1_generator_mul_run(d_scalars, n, d_out);
}
int cqb_graph_evaluate_dev(const cqb_graph_t* graph, const void* const* d_fixed, uint32_t n_fixed, const void* const* d_advice,
                           uint32_t n_advice, const void* const* d_instance, uint32_t n_instance, const uint64_t* challenges,
                           uint32_t n_challenges, const uint64_t beta[4], const uint64_t gamma[4], const uint64_t theta[4],
                           const uint64_t y[4], void* d_values, uint64_t size, int32_t rot_scale) {
    LOCK;
    CQB_TRY(require_init());
    if (!graph || !beta || !gamma || !theta || !y || (!d_values && size)) return fail(CQB_E_BAD_ARG, "cqb_graph_evaluate_dev: NULL argument");
    return graph_evaluate_run(graph, d_fixed, n_fixed, d_advice, n_advice, d_instance, n_instance, challenges, n_challenges, beta, gamma, theta,
                              y, d_values, size, rot_scale);
}
int cqb_cq_lookup_h_dev(void* d_values, const void* d_b_coset, const void* d_f_coset, const void* d_l_active_row, const uint64_t beta[4],
                        const uint64_t y[4], uint64_t size) {
    LOCK;
    CQB_TRY(require_init());
    if (!beta || !y || ((!d_values || !d_b_coset || !d_f_coset || !d_l_active_row) && size)) return fail(CQB_E_BAD_ARG, "cqb_cq_lookup_h_dev: NULL argument");
    return cq_lookup_h_run(d_values, d_b_coset, d_f_coset, d_l_active_row, beta, y, size);
}
int cqb_permutation_h_dev(void* d_values, uint64_t size, int32_t rot_scale, int32_t last_rotation, uint32_t chunk_len,
                          const void* const* d_sets, uint32_t nsets, const void* const* d_columns, const void* const* d_perm_cosets,
                          uint32_t ncols, const void* d_l0, const void* d_l_last, const void* d_l_active_row, const uint64_t beta[4],
                          const uint64_t gamma[4], const uint64_t y[4], const uint64_t extended_omega[4]) {
    LOCK;
    CQB_TRY(require_init());
    if (!beta || !gamma || !y || !extended_omega) return fail(CQB_E_BAD_ARG, "cqb_permutation_h_dev: NULL argument");
    return permutation_h_run(d_values, size, rot_scale, last_rotation, chunk_len, d_sets, nsets, d_columns, d_perm_cosets, ncols, d_l0, d_l_last,
                             d_l_active_row, beta, gamma, y, extended_omega);
}
int cqb_eval_polynomial_dev(const void* d_coeffs, size_t n, const uint64_t point[4], uint64_t out[4]) {
    LOCK;
    CQB_TRY(require_init());
    if ((!d_coeffs && n) || !point || !out) return fail(CQB_E_BAD_ARG, "cqb_eval_polynomial_dev: NULL argument");
    CQB_TRY(eval_polynomial_run(d_coeffs, n, point, g_out->p));
    CQB_TRY(g_out_host->ensure(128));
    CQB_CUDA(cudaMemcpyAsync(g_out_host->p, g_out->p, 32, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    memcpy(out, g_out_host->p, 32);
    return 0;
}
// `count` evaluations queued back to back, ONE read-back: create_proof evaluates every queried polynomial at x, omega x, ... between two
// challenges (plonk/prover.rs:629-719), and a 32-byte synchronous read per evaluation costs more than the evaluation at circuit sizes
int cqb_eval_polynomials_dev(const void* const* d_coeffs, size_t n, const uint64_t* points, uint32_t count, uint64_t* out) {
    LOCK;
    CQB_TRY(require_init());
    if (count == 0) return 0;
    if (!d_coeffs || !points || !out) return fail(CQB_E_BAD_ARG, "cqb_eval_polynomials_dev: NULL argument");
    CQB_TRY(g_out->ensure((size_t)count * 32 + 256));
    CQB_TRY(g_out_host->ensure((size_t)count * 32 + 256));
    for (uint32_t i = 0; i < count; i++) {
        if (!d_coeffs[i] && n) return fail(CQB_E_BAD_ARG, "cqb_eval_polynomials_dev: NULL polynomial %u", i);
        CQB_TRY(eval_polynomial_run(d_coeffs[i], n, points + 4 * i, (char*)g_out->p + (size_t)i * 32));
    }
    CQB_CUDA(cudaMemcpyAsync(g_out_host->p, g_out->p, (size_t)count * 32, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    memcpy(out, g_out_host->p, (size_t)count * 32);
    return 0;
}
int cqb_kate_division_dev(const void* d_a, size_t n, const uint64_t b[4], void* d_q) {
    LOCK;
    CQB_TRY(require_init());
    if (!b || ((!d_a || !d_q) && n > 1) || (d_a == d_q && n > 1)) return fail(CQB_E_BAD_ARG, "cqb_kate_division_dev: NULL or aliasing arguments");
    return kate_division_run(d_a, n, b, d_q);
}
int cqb_fr_prefix_product_dev(const void* d_in, size_t n, const uint64_t init[4], void* d_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!init || ((!d_in || !d_out) && n)) return fail(CQB_E_BAD_ARG, "cqb_fr_prefix_product_dev: NULL argument");
    return fr_prefix_product_run(d_in, n, init, d_out);
}
int cqb_permutation_product_dev(const void* const* d_columns, const void* const* d_perms, uint32_t ncols, uint32_t k, const uint64_t beta[4],
                                const uint64_t gamma[4], const uint64_t omega[4], const uint64_t delta[4], uint64_t deltaomega_io[4],
                                const uint64_t last_z[4], void* d_z) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_columns || !d_perms || !beta || !gamma || !omega || !delta || !deltaomega_io || !last_z || !d_z)
        return fail(CQB_E_BAD_ARG, "cqb_permutation_product_dev: NULL argument");
    for (uint32_t j = 0; j < ncols; j++)
        if (!d_columns[j] || !d_perms[j]) return fail(CQB_E_BAD_ARG, "cqb_permutation_product_dev: NULL column %u", j);
    return permutation_product_run(d_columns, d_perms, ncols, k, beta, gamma, omega, delta, deltaomega_io, last_z, d_z);
}
int cqb_fr_compress_dev(const void* const* d_cols, uint32_t ncols, const uint32_t* d_idx, size_t n, const uint64_t theta[4], void* d_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_cols || !theta || (!d_out && n)) return fail(CQB_E_BAD_ARG, "cqb_fr_compress_dev: NULL argument");
    for (uint32_t k = 0; k < ncols; k++)
        if (!d_cols[k] && n) return fail(CQB_E_BAD_ARG, "cqb_fr_compress_dev: NULL column %u", k);
    return fr_compress_run(d_cols, ncols, d_idx, n, theta, d_out);
}
int cqb_fr_inv_shifted_dev(const void* d_in, size_t n, size_t usable, const uint64_t shift[4], void* d_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!shift || ((!d_in || !d_out) && n)) return fail(CQB_E_BAD_ARG, "cqb_fr_inv_shifted_dev: NULL argument");
    if (usable > n) return fail(CQB_E_LEN_MISMATCH, "cqb_fr_inv_shifted_dev: usable rows %zu exceed n = %zu", usable, n);
    return fr_inv_shifted_run(d_in, n, usable, shift, d_out);
}
int cqb_fr_mul_dev(const void* d_a, const void* d_b, size_t n, void* d_out) {
    LOCK;
    CQB_TRY(require_init());
    if ((!d_a || !d_b || !d_out) && n) return fail(CQB_E_BAD_ARG, "cqb_fr_mul_dev: NULL argument");
    return fr_mul_run(d_a, d_b, n, d_out);
}
int cqb_fr_axpy_dev(void* d_acc, const uint64_t a[4], const void* d_x, size_t n) {
    LOCK;
    CQB_TRY(require_init());
    if ((!d_acc || !d_x || !a) && n) return fail(CQB_E_BAD_ARG, "cqb_fr_axpy_dev: NULL argument");
    return fr_axpy_run(d_acc, a, d_x, n);
}
int cqb_msm_bn254_g1_sparse_dev(cqb_bases_t b, const uint32_t* d_idx, const void* d_scalars, size_t m, uint64_t out_xy[8], int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || ((!d_scalars || !d_idx) && m)) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g1_sparse_dev: NULL argument");
    auto it = g_bases.find(b);
    if (it == g_bases.end()) return fail(CQB_E_BAD_ARG, "unknown bases handle %llu", (unsigned long long)b);
    CQB_TRY(dispatch_msm(&it->second, 0, d_scalars, d_idx, m));  // indices are the caller's responsibility (device-resident)
    return fetch_result(out_xy, is_inf);
}
int cqb_lookup_product_dev(const void* d_compressed_input, const void* d_compressed_table, const void* d_permuted_input,
                           const void* d_permuted_table, uint32_t k, const uint64_t beta[4], const uint64_t gamma[4], void* d_z) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_compressed_input || !d_compressed_table || !d_permuted_input || !d_permuted_table || !beta || !gamma || !d_z)
        return fail(CQB_E_BAD_ARG, "cqb_lookup_product_dev: NULL argument");
    return lookup_product_run(d_compressed_input, d_compressed_table, d_permuted_input, d_permuted_table, k, beta, gamma, d_z);
}
int cqb_lookup_h_dev(void* d_values, const void* d_table_value, const void* d_product_coset, const void* d_permuted_input_coset,
                     const void* d_permuted_table_coset, const void* d_l0, const void* d_l_last, const void* d_l_active_row, const uint64_t beta[4],
                     const uint64_t gamma[4], const uint64_t y[4], uint64_t size, int32_t rot_scale) {
    LOCK;
    CQB_TRY(require_init());
    if (!beta || !gamma || !y || ((!d_values || !d_table_value || !d_product_coset || !d_permuted_input_coset || !d_permuted_table_coset || !d_l0 ||
                                   !d_l_last || !d_l_active_row) && size))
        return fail(CQB_E_BAD_ARG, "cqb_lookup_h_dev: NULL argument");
    return lookup_h_run(d_values, d_table_value, d_product_coset, d_permuted_input_coset, d_permuted_table_coset, d_l0, d_l_last, d_l_active_row, beta,
                        gamma, y, size, rot_scale);
}
static Scratch g_scale_tab;
int cqb_fr_scale_dev(void* d_a, size_t n, const uint64_t factor[4]) {
    LOCK;
    CQB_TRY(require_init());
    if (!factor || (!d_a && n)) return fail(CQB_E_BAD_ARG, "cqb_fr_scale_dev: NULL argument");
    if (n == 0) return 0;
    CQB_TRY(g_scale_tab.ensure(64));
    CQB_CUDA(cudaMemcpyAsync(g_scale_tab.p, factor, 32, cudaMemcpyHostToDevice, g_ctx.stream));
    return fr_scale_table(d_a, n, g_scale_tab.p, 1);
}
int cqb_fr_batch_invert_dev(void* d_a, size_t n) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_a && n) return fail(CQB_E_BAD_ARG, "cqb_fr_batch_invert_dev: NULL argument");
    return fr_batch_invert_run(d_a, n);
}
int cqb_fr_powers_dev(const uint64_t base[4], size_t n, void* d_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!base || (!d_out && n)) return fail(CQB_E_BAD_ARG, "cqb_fr_powers_dev: NULL argument");
    return fr_powers_run(base, n, d_out);
}

// ---- memory helpers ---------------------------------------------------------------------------------------------
int cqb_dev_alloc(size_t bytes, void** d_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!d_out) return fail(CQB_E_BAD_ARG, "cqb_dev_alloc: NULL argument");
    if (cudaMalloc(d_out, bytes ? bytes : 32) != cudaSuccess) {
        cudaGetLastError();
        msm_release_scratch();  // the MSM's grow-only working buffers (up to ~20 GB after a 2^24+ MSM) make room; the next MSM reallocates
        if (cudaMalloc(d_out, bytes ? bytes : 32) != cudaSuccess) {
            cudaGetLastError();
            return fail(CQB_E_OOM, "cudaMalloc(%zu) failed", bytes);
        }
    }
    return 0;
}
int cqb_dev_free(void* d) {
    LOCK;
    CQB_TRY(require_init());
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    CQB_CUDA(cudaFree(d));
    return 0;
}
int cqb_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes) {
    LOCK;
    CQB_TRY(require_init());
    CQB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return 0;
}
int cqb_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes) {
    LOCK;
    CQB_TRY(require_init());
    CQB_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, g_ctx.stream));
    return 0;
}
int cqb_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes) {
    LOCK;
    CQB_TRY(require_init());
    CQB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return 0;
}
int cqb_host_alloc_pinned(size_t bytes, void** h_out) {
    LOCK;
    CQB_TRY(require_init());
    if (!h_out) return fail(CQB_E_BAD_ARG, "cqb_host_alloc_pinned: NULL argument");
    if (cudaMallocHost(h_out, bytes ? bytes : 32) != cudaSuccess) {
        cudaGetLastError();
        return fail(CQB_E_OOM, "cudaMallocHost(%zu) failed", bytes);
    }
    return 0;
}
int cqb_host_free_pinned(void* h) {
    LOCK;
    CQB_CUDA(cudaFreeHost(h));
    return 0;
}

int cqb_msm_set_profiling(int on) {
    LOCK;
    msm_set_profiling(on != 0);
    return 0;
}
int cqb_msm_phase_ms(float* ms, int cap) {
    LOCK;
    if (!ms) return 0;
    return msm_phase_ms(ms, cap);
}

static int check_slot(int slot) {
    if (slot < 0 || slot >= g_nslots) return fail(CQB_E_BAD_ARG, "device slot %d out of range (%d active)", slot, g_nslots);
    return 0;
}
int cqb_dev_alloc_on(int slot, size_t bytes, void** d_out) {
    LOCK;
    CQB_TRY(require_init());
    CQB_TRY(check_slot(slot));
    if (!d_out) return fail(CQB_E_BAD_ARG, "cqb_dev_alloc_on: NULL argument");
    bind_slot(slot);
    cudaError_t e = cudaMalloc(d_out, bytes ? bytes : 64);
    if (e != cudaSuccess) {
        cudaGetLastError();
        msm_release_scratch();
        e = cudaMalloc(d_out, bytes ? bytes : 64);
    }
    bind_slot(0);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(CQB_E_OOM, "cudaMalloc(%zu) on slot %d failed", bytes, slot); }
    return 0;
}
int cqb_dev_free_on(int slot, void* d) {
    LOCK;
    CQB_TRY(check_slot(slot));
    bind_slot(slot);
    cudaDeviceSynchronize();
    cudaError_t e = cudaFree(d);
    bind_slot(0);
    if (e != cudaSuccess) return fail(CQB_E_CUDA, "cudaFree on slot %d failed: %s", slot, cudaGetErrorString(e));
    return 0;
}
int cqb_memcpy_h2d_on(int slot, void* d_dst, const void* h_src, size_t bytes) {
    LOCK;
    CQB_TRY(check_slot(slot));
    bind_slot(slot);
    cudaError_t e = cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice);
    bind_slot(0);
    if (e != cudaSuccess) return fail(CQB_E_CUDA, "cudaMemcpy to slot %d failed: %s", slot, cudaGetErrorString(e));
    return 0;
}
// ---- G2: the verifier-side half of the SRS and the CQ table commitment (keygen-time, primary device) ----------------------
int cqb_g2_generator_mul_dev(const void* d_scalars, size_t n, void* d_out_affine) {
    LOCK;
    CQB_TRY(require_init());
    if ((!d_scalars || !d_out_affine) && n) return fail(CQB_E_BAD_ARG, "cqb_g2_generator_mul_dev: NULL argument");
    return g2_mul_run(nullptr, d_scalars, n, d_out_affine);
}
int cqb_g2_powers(const uint64_t s[4], size_t count, uint64_t* g2_affine_out) {
    LOCK;
    CQB_TRY(require_init());
    if ((!s || !g2_affine_out) && count) return fail(CQB_E_BAD_ARG, "cqb_g2_powers: NULL argument");
    if (count == 0) return 0;
    CQB_TRY(g_io.ensure(count * 32 + count * 128));
    void* d_pw = g_io.p;
    void* d_out = (char*)g_io.p + count * 32;
    CQB_TRY(fr_powers_run(s, count, d_pw));
    CQB_TRY(g2_mul_run(nullptr, d_pw, count, d_out));
    CQB_CUDA(cudaMemcpyAsync(g2_affine_out, d_out, count * 128, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return 0;
}
int cqb_msm_bn254_g2(const uint64_t* g2_affine, const uint64_t* scalars, size_t n, uint64_t out_xy[16], int* is_inf) {
    LOCK;
    CQB_TRY(require_init());
    if (!out_xy || ((!g2_affine || !scalars) && n)) return fail(CQB_E_BAD_ARG, "cqb_msm_bn254_g2: NULL argument");
    CQB_TRY(g_io.ensure(n * 160 + 256));
    char* d_b = (char*)g_io.p + 256;
    char* d_s = d_b + n * 128;
    if (n) {
        CQB_CUDA(cudaMemcpyAsync(d_b, g2_affine, n * 128, cudaMemcpyHostToDevice, g_ctx.stream));
        CQB_CUDA(cudaMemcpyAsync(d_s, scalars, n * 32, cudaMemcpyHostToDevice, g_ctx.stream));
    }
    CQB_TRY(g2_msm_run(d_b, d_s, n, g_io.p));
    CQB_TRY(g_out_host->ensure(256));
    CQB_CUDA(cudaMemcpyAsync(g_out_host->p, g_io.p, 144, cudaMemcpyDeviceToHost, g_ctx.stream));
    CQB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    memcpy(out_xy, g_out_host->p, 128);
    if (is_inf) *is_inf = (int)((uint32_t*)g_out_host->p)[32];
    return 0;
}

int cqb_msm_set_window_bits(int c) {
    LOCK;
    if (c != 0 && (c < 2 || c > 16)) return fail(CQB_E_BAD_ARG, "window bits must be 0 (auto) or 2..16");
    msm_set_window_bits(c);
    return 0;
}

int cqb_msm_set_accumulator(int mode, int affine_seg_log) {
    LOCK;
    if (mode < 0 || mode > 3 || affine_seg_log < 0 || affine_seg_log > 10) return fail(CQB_E_BAD_ARG, "cqb_msm_set_accumulator: mode 0..3, segment log 0..10");
    msm_set_accumulator(mode);
    msm_set_affine_segment(affine_seg_log);
    if (const char* v = getenv("CQB_AFF_VARIANT")) msm_set_affine_variant(atoi(v));
    if (const char* v = getenv("CQB_TREE_LEVELS")) msm_set_tree_levels(atoi(v));
    if (const char* v = getenv("CQB_SORT")) msm_set_sort_mode(atoi(v));
    if (const char* v = getenv("CQB_TREE_SLABS")) msm_set_tree_slabs(atoi(v));
    return 0;
}
int cqb_msm_set_sort_mode(int mode) {
    LOCK;
    if (mode < 0 || mode > 2) return fail(CQB_E_BAD_ARG, "cqb_msm_set_sort_mode: 0..2");
    msm_set_sort_mode(mode);
    return 0;
}
int cqb_msm_last_tree_levels(void) { return msm_last_tree_levels(); }
int cqb_msm_set_tree_levels(int levels) {
    LOCK;
    if (levels < 1 || levels > 6) return fail(CQB_E_BAD_ARG, "cqb_msm_set_tree_levels: 1..6");
    msm_set_tree_levels(levels);
    return 0;
}
int cqb_msm_set_parts(int parts) {
    LOCK;
    if (parts < 0 || parts > 8) return fail(CQB_E_BAD_ARG, "parts must be 0 (auto) or 1..8");
    msm_set_parts(parts);
    return 0;
}

}  // extern "C"

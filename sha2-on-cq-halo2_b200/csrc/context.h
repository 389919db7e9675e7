// context.h — host-side runtime shared by the kernels' launchers: one context per process (one process per GPU),
// error plumbing for the C ABI (every entry point returns 0 or a CQB_E_* code; the message is kept for cqb_last_error),
// a stream (the library's own, or the caller's — e.g. torch's current stream — via cqb_set_stream) and grow-only scratch
// buffers so that steady-state calls perform no cudaMalloc.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/cqb200.h"

namespace cqb {

// One context per DEVICE SLOT. A process drives one GPU (cqb_init(device): slot 0 only — the one-process-per-GPU form the
// benchmark launches under torchrun) or several (cqb_init_multi(n): slot i = device i; MSMs over a sharded base set fan out
// to one host thread per slot, the reference's own decomposition across threads, arithmetic.rs:137-153). Every host thread
// works on the slot it is bound to (thread-local); ctx() and every PerDevice<> object resolve through it, so the launchers
// below the ABI are written once and never mention a device.
constexpr int MAX_DEVICES = 16;
int cur_slot();
void bind_slot(int slot);  // also cudaSetDevice()s the slot's device when the slot is initialised

template <class T>
struct PerDevice {
    T v[MAX_DEVICES];
    T* operator->() { return &v[cur_slot()]; }
    T& get() { return v[cur_slot()]; }
    T& at(int slot) { return v[slot]; }
};

struct Ctx {
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool inited = false;
    std::string last_error;
    unsigned long long launches = 0;  // kernels launched by this library since init / last reset (bench's gpu_launches)
};
Ctx& ctx();

int fail(int code, const char* fmt, ...);

#define CQB_CUDA(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) return ::cqb::fail(CQB_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, \
                                                  cudaGetErrorString(_e));                                 \
    } while (0)

#define CQB_TRY(expr)          \
    do {                       \
        int _rc = (expr);      \
        if (_rc != 0) return _rc; \
    } while (0)

#define CQB_LAUNCHED() (::cqb::ctx().launches++)

// grow-only device scratch buffer
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// pinned host staging buffer (grow-only)
struct Pinned {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

}  // namespace cqb

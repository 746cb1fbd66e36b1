// Shared helpers for the sm_100a kernels of the DSSM hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/dssm_b200.h"

namespace dssm {

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
// counts kernel launches (bench.py's gpu_launches); bumped by LAUNCH_CHECK
extern thread_local int64_t g_launch_count;

#define DSSM_REQUIRE(cond, code, ...)                 \
    do {                                              \
        if (!(cond)) return ::dssm::fail((code), __VA_ARGS__); \
    } while (0)

#define CUDA_TRY(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return ::dssm::fail(DSSM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                __FILE__, __LINE__);                                               \
    } while (0)

// after every kernel launch: no host sync, only the launch error
#define LAUNCH_CHECK(name)                                                                         \
    do {                                                                                           \
        ::dssm::g_launch_count++;                                                                  \
        cudaError_t _e = cudaGetLastError();                                                       \
        if (_e != cudaSuccess)                                                                     \
            return ::dssm::fail(DSSM_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
int sm_count();

// "do once per device" latch for cudaFuncSetAttribute calls (function attributes are per device, and a process may
// drive several devices): `static PerDeviceOnce once; if (once.need()) { ...set attributes... }`
struct PerDeviceOnce {
    bool done[64] = {};
    bool need() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

// bump allocator over a caller-provided workspace
struct Arena {
    char* base;
    size_t cap, off;
    Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= cap; }
};

// ---- device helpers -----------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float act_fwd(float x, int act) {
    if (act == DSSM_ACT_RELU) return fmaxf(x, 0.f);
    if (act == DSSM_ACT_TANH) return tanhf(x);
    return x;
}
// derivative expressed through the activation output a
__device__ __forceinline__ float act_grad_from_out(float a, int act) {
    if (act == DSSM_ACT_RELU) return a > 0.f ? 1.f : 0.f;
    if (act == DSSM_ACT_TANH) return 1.f - a * a;
    return 1.f;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// streaming 128-bit load that does not pollute L1 (read-once data)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
// vector reduction to global memory (sm_90+): no return value, 16 bytes per request
__device__ __forceinline__ void red_add4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
#endif

}  // namespace dssm

// float4 chunks per lane for a row of L floats handled by one warp: compile-time unrolled for 1..8 (L <= 1024)
#define DISPATCH_NCH(nch, CALL)                       \
    switch (nch) {                                    \
        case 1: { constexpr int N_ = 1; CALL; } break; \
        case 2: { constexpr int N_ = 2; CALL; } break; \
        case 3: { constexpr int N_ = 3; CALL; } break; \
        case 4: { constexpr int N_ = 4; CALL; } break; \
        case 5: { constexpr int N_ = 5; CALL; } break; \
        case 6: { constexpr int N_ = 6; CALL; } break; \
        case 7: { constexpr int N_ = 7; CALL; } break; \
        default: { constexpr int N_ = 8; CALL; } break; \
    }

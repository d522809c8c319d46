// Shared helpers for libnbe_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <utility>
#include "../../include/nbe_b200.h"

namespace nbe {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// Call after every kernel launch: counts it and converts launch errors to a status.
inline int launched(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(NBE_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    return NBE_OK;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// The batch step is a chain of ~35 kernels, most of them persistent with a per-launch prologue (barrier init, TMEM allocation,
// cluster sync, tensor-map prefetch).  Launched with cudaLaunchAttributeProgrammaticStreamSerialization, a
// kernel's CTAs may become resident while the tail of its predecessor is still running: every kernel of the chain calls
// pdl_trigger() first (its dependents may start launching once all of its CTAs have started) and pdl_wait() after its prologue,
// BEFORE the first access to global memory another kernel may have written or may still be reading (the wait returns when every
// prerequisite grid has completed and its memory is visible).  Both are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();                                                // NBE_NO_PDL=1: plain stream-ordered launches (A/B switch)

template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);       // errors surface through launched()
}

#define NBE_REQUIRE(cond, ...) do { if (!(cond)) return nbe::fail(NBE_EINVAL, __VA_ARGS__); } while (0)

constexpr int kNumSMs = 148;   // B200

template <class T> struct Cvt;
template <> struct Cvt<float> {
    static __device__ __forceinline__ float ld(float v) { return v; }
    static __device__ __forceinline__ float st(float v) { return v; }
};
template <> struct Cvt<__half> {
    static __device__ __forceinline__ float ld(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half st(float v) { return __float2half_rn(v); }
};
template <> struct Cvt<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 st(float v) { return __float2bfloat16_rn(v); }
};

// 16-byte streaming accesses (no L1 allocation: every byte is touched once).
__device__ __forceinline__ int4 ld_stream16(const void* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void* p, const int4& v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Packed FP32 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 — two IEEE-rn lanes per issued instruction).  The scalar FP32 pipe
// issues one 3-register FFMA per two cycles per SM sub-partition; the chip's FP32 peak is only reachable through these.
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
          "l"(*reinterpret_cast<const unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
// two bf16 in one 32-bit word -> two exact floats (low half first)
__device__ __forceinline__ float2 bf16x2_to_f2(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

__device__ __forceinline__ float lrelu_gain_clamp(float v, float alpha, float gain, float clamp) {
    v = (v > 0.f) ? v : v * alpha;
    v *= gain;
    if (clamp >= 0.f) v = fminf(fmaxf(v, -clamp), clamp);
    return v;
}

// Generic forward activation in fp32 (semantics of bias_act.cu:39-130 / bias_act.py:23-33).
__device__ __forceinline__ float apply_act(float x, int act, float alpha) {
    switch (act) {
        default:
        case NBE_ACT_LINEAR:   return x;
        case NBE_ACT_RELU:     return x > 0.f ? x : 0.f;
        case NBE_ACT_LRELU:    return x > 0.f ? x : x * alpha;
        case NBE_ACT_TANH:     return tanhf(x);
        case NBE_ACT_SIGMOID:  return 1.f / (1.f + expf(-x));
        case NBE_ACT_ELU:      return x >= 0.f ? x : expm1f(x);
        case NBE_ACT_SELU:     return 1.0507009873554804934193349852946f *
                                      (x >= 0.f ? x : 1.6732632423543772848170429916717f * expm1f(x));
        case NBE_ACT_SOFTPLUS: return x > 20.f ? x : log1pf(expf(x));
        case NBE_ACT_SWISH:    return x / (1.f + expf(-x));
    }
}

}  // namespace nbe

// Shared helpers for the plagnn sm_100a kernels (internal; the public surface is include/plagnn.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/plagnn.h"

namespace plagnn {

// Thread-local last-error text returned by plagnn_last_error().
void set_error(const char* fmt, ...);

inline int fail(int code, const char* what, const char* detail = "") {
    set_error("%s%s%s", what, detail[0] ? ": " : "", detail);
    return code;
}

// Host-side count of kernels launched by this library (bench.py's "gpu_launches").
void count_launches(int n);

// Checks the launch (not the execution): entry points never synchronise the stream.
inline int check_launch(const char* what, int launched = 1) {
    count_launches(launched);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: %s", what, cudaGetErrorString(e));
        return PLAGNN_ERR_CUDA;
    }
    return PLAGNN_OK;
}

#define PLAGNN_CUDA_TRY(expr)                                                      \
    do {                                                                           \
        cudaError_t _e = (expr);                                                   \
        if (_e != cudaSuccess) {                                                   \
            plagnn::set_error("%s: %s", #expr, cudaGetErrorString(_e));           \
            return PLAGNN_ERR_CUDA;                                                \
        }                                                                          \
    } while (0)

// Optional per-call CUDA-event timing on the launching stream (plagnn_profile_enable / _report): lets bench.py
// measure each kernel's duration inside its own timed region without a profiler attached.
struct ProfileScope {
    int slot;
    cudaStream_t st;
    ProfileScope(const char* name, long long t0, long long t1, long long t2, plagnn_stream_t stream);
    ~ProfileScope();
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();   // cached cudaDevAttrMultiProcessorCount of the current device

// Programmatic dependent launch: every kernel of the epoch is launched with the stream-serialisation attribute relaxed, so
// the next grid's CTAs are placed (and run their set-up: barrier init, TMEM allocation, descriptor prefetch) while the
// previous grid drains.  Kernels call pdl_trigger() on entry and pdl_wait() before their first access to global memory;
// pdl_wait() returns once the preceding grid has completed and its writes are visible.  PLAGNN_PDL=0 restores plain
// stream order (the device-side instructions are then no-ops).
bool pdl_enabled();

template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

// ---- device helpers ------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// entry sequence of a kernel without set-up work worth overlapping
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    switch (act) {
        case PLAGNN_ACT_RELU:    return v > 0.f ? v : 0.f;
        case PLAGNN_ACT_LEAKY:   return v > 0.f ? v : v * slope;
        case PLAGNN_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default:                 return v;
    }
}

// derivative of the activation expressed through the SAVED FORWARD OUTPUT y = act(z)
__device__ __forceinline__ float act_grad_from_output(float y, int act, float slope) {
    switch (act) {
        case PLAGNN_ACT_RELU:    return y > 0.f ? 1.f : 0.f;
        case PLAGNN_ACT_LEAKY:   return y > 0.f ? 1.f : slope;
        case PLAGNN_ACT_SIGMOID: return y * (1.f - y);
        default:                 return 1.f;
    }
}

}  // namespace plagnn

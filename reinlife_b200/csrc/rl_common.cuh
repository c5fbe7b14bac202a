// rl_common.cuh -- shared device helpers for the ReinLife sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/reinlife_b200.h"

#define RL_NONE16 0xFFFFu

extern thread_local char g_rl_err[512];
int rl_set_err(int code, const char* fmt, ...);

// sums the per-CTA gradient slabs (fixed order) into learn->grad; grad[n_train] = *ev_total (learn_kernels.cu)
// w1_rowmajor = 1: the slabs hold the first-layer gradient as [N1][160] (tensor-core path) instead of W1t [160][N1]
int rl_learn_reduce(const rl_learn_bufs* learn, const int32_t* ev_total, int w1_rowmajor, void* stream);

#define RL_CUDA_CHECK(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return rl_set_err(RL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                            \
    } while (0)

// One-time setup per DEVICE (cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are per-device settings;
// a process may drive several GPUs through Environment(device=...)).
struct PerDeviceOnce {
    bool done[64] = {};
    bool need() {
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};
inline int rl_device_sm_count() {
    static int sm[64] = {};
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (!sm[d]) {
        cudaDeviceGetAttribute(&sm[d], cudaDevAttrMultiProcessorCount, d);
        if (sm[d] <= 0) sm[d] = 148;
    }
    return sm[d];
}

#define RL_ARG_CHECK(cond)                                                                    \
    do {                                                                                      \
        if (!(cond)) return rl_set_err(RL_ERR_ARG, "argument check failed: %s (%s:%d)", #cond, __FILE__, __LINE__); \
    } while (0)

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streaming (write-once) stores: keep obs rows from polluting L1
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {      // read-once data: do not keep it in L1
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream_i4(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

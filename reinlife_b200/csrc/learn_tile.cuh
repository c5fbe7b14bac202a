// learn_tile.cuh -- fp32 building blocks shared by the learn kernels (learn_kernels.cu: dueling events;
// learn_rows_kernels.cu: DQN and PPO row tiles): block reductions, weight-gradient outer products into the CTA-private
// slab, column sums, replay-row gathers and the small head GEMV.
#pragma once
#include <cuda_fp16.h>
#include "mlp_tile.cuh"
#include "models.cuh"

namespace mlp {

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane_id() == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) r += red[i];
    __syncthreads();
    return r;
}

// G[m][n] += sum_b A[b][m] * B[b][n]   (b < 64; A, B in shared memory; G = CTA-private global slab, row-major, leading dim ldg)
template <int M, int N, int TM, int TN>
__device__ __forceinline__ void outer_accum(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                            float* __restrict__ G, int ldg = N) {
    constexpr int NG = TN / 4, GS = N / NG;
    static_assert((M / TM) * (N / TN) == NT && TM % 4 == 0 && TN % 4 == 0, "thread tiling");
    const int tx = threadIdx.x % (N / TN), ty = threadIdx.x / (N / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int b = 0; b < R; ++b) {
        float av[TM], bv[TN];
#pragma unroll
        for (int i4 = 0; i4 < TM / 4; ++i4) {
            const float4 t = *reinterpret_cast<const float4*>(A + (size_t)b * lda + ty * TM + i4 * 4);
            av[i4 * 4 + 0] = t.x; av[i4 * 4 + 1] = t.y; av[i4 * 4 + 2] = t.z; av[i4 * 4 + 3] = t.w;
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float4 t = *reinterpret_cast<const float4*>(B + (size_t)b * ldb + g * GS + tx * 4);
            bv[g * 4 + 0] = t.x; bv[g * 4 + 1] = t.y; bv[g * 4 + 2] = t.z; bv[g * 4 + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            float4* p = reinterpret_cast<float4*>(G + (size_t)(ty * TM + i) * ldg + g * GS + tx * 4);
            float4 o = *p;
            o.x += acc[i][g * 4 + 0]; o.y += acc[i][g * 4 + 1]; o.z += acc[i][g * 4 + 2]; o.w += acc[i][g * 4 + 3];
            *p = o;
        }
}

// G[n] += sum_b B[b][n]
template <int N>
__device__ __forceinline__ void colsum_accum(const float* __restrict__ B, int ldb, float* __restrict__ G) {
    for (int n = threadIdx.x; n < N; n += NT) {
        float s = 0.f;
#pragma unroll 8
        for (int b = 0; b < R; ++b) s += B[(size_t)b * ldb + n];
        G[n] += s;
    }
}

// the same gather from a float16 ring (rl_replay_bufs.obs_fp16): fp32 arithmetic on the rows as they were rounded at store time
__device__ __forceinline__ void gather64_h(float* dst, int ldd, const void* __restrict__ src, const int* ids) {
    const uint2* s2 = reinterpret_cast<const uint2*>(src);
    for (int v = threadIdx.x; v < R * (RL_K1 / 4); v += NT) {
        const int r = v / (RL_K1 / 4), c4 = v - r * (RL_K1 / 4);
        const uint2 h = __ldg(s2 + (size_t)ids[r] * (RL_K1 / 4) + c4);
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
        *reinterpret_cast<float4*>(dst + (size_t)r * ldd + c4 * 4) = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
}

__device__ __forceinline__ void gather64(float* dst, int ldd, const float* __restrict__ src, const int* ids) {
    for (int v = threadIdx.x; v < R * (RL_K1 / 4); v += NT) {
        const int r = v / (RL_K1 / 4), c4 = v - r * (RL_K1 / 4);
        *reinterpret_cast<float4*>(dst + (size_t)r * ldd + c4 * 4) =
            __ldg(reinterpret_cast<const float4*>(src + (size_t)ids[r] * RL_K1) + c4);
    }
}

template <int N2, int NH>
__device__ __forceinline__ void head64(const float* H2, int ldh, const float* Wh_s, float* OUT) {
    for (int o = threadIdx.x; o < R * NH; o += NT) {
        const int r = o / NH, j = o - r * NH;
        const float* h = H2 + (size_t)r * ldh;
        float acc = Wh_s[N2 * NH + j];
#pragma unroll 8
        for (int k = 0; k < N2; ++k) acc = fmaf(h[k], Wh_s[k * NH + j], acc);
        OUT[r * 16 + j] = acc;
    }
    __syncthreads();
}

}  // namespace mlp

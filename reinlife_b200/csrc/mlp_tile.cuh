// mlp_tile.cuh -- fp32 tile GEMM stage shared by the brain forward and learn kernels.
//
// One CTA (256 threads) owns a tile of R=64 rows whose activations stay in shared memory for the whole
// network; weights (k-major, L2-resident, <= 216 KB per network) are streamed through shared memory in
// 16 KB chunks by the bulk async-copy engine (cp.async.bulk + mbarrier, SASS UBLKCP) double-buffered
// against the FFMA loop.  fp32 FMA is used on purpose: parity with the reference's fp32 torch path is
// stated at rtol 1e-4 (tests/test_brain_gpu.py).
#pragma once
#include "rl_common.cuh"

namespace mlp {

constexpr int R = 64;            // rows per tile (= PERD3QN batch_size: one train() event per tile)
constexpr int NT = 256;          // threads per CTA
constexpr int CHUNK_BYTES = 16384;

template <int N> struct Cfg;
template <> struct Cfg<64>  { static constexpr int TM = 4, TN = 4; };
template <> struct Cfg<128> { static constexpr int TM = 4, TN = 8; };
template <> struct Cfg<256> { static constexpr int TM = 8, TN = 8; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    fence_proxy_async();   // order earlier generic-proxy reads of dst before the async-proxy overwrite
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(bar);
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}

// weight-chunk pipeline state shared by consecutive stages of one CTA
struct Pipe {
    float* wbuf;        // [2][CHUNK_BYTES/4]
    uint64_t* bars;     // [2]
    uint32_t phase;     // bit b = parity to wait for on barrier b
};

// OUT[r][n] = act( bias[n] + sum_k A[r][k] * Wg[k][n] ),  r < 64, n < N, k < K
//   A   : shared, row-major, leading dim lda (floats, multiple of 4)
//   Wg  : global, k-major [K][N] contiguous, 16-byte aligned
//   OUT : shared, row-major, leading dim ldo; must not alias A
//   MODE: 0 plain, 1 ReLU, 2 backward mask -- OUT[r][n] = OUT_old[r][n] > 0 ? value : 0 (ReLU derivative, in place)
// Each thread owns TM rows x TN columns, the columns interleaved in groups of 4 (col = g*(N/NG) + tx*4 + jj) so
// that every 128-bit shared load of a weight row is a contiguous, conflict-free 16 B x 32 lanes access.
template <int K, int N, int MODE, bool HAS_BIAS>
__device__ __forceinline__ void gemm_stage(const float* __restrict__ A, int lda, const float* __restrict__ Wg,
                                           const float* __restrict__ bias, float* OUT, int ldo, Pipe& pp) {
    constexpr int TM = Cfg<N>::TM, TN = Cfg<N>::TN, NG = TN / 4, GS = N / NG;
    constexpr int KC = CHUNK_BYTES / (4 * N);
    static_assert(K % KC == 0 && KC % 4 == 0, "K must be a multiple of the chunk depth");
    static_assert((R / TM) * (N / TN) == NT && TN % 4 == 0, "thread tiling");
    constexpr int NCH = K / KC;
    const int tx = threadIdx.x % (N / TN), ty = threadIdx.x / (N / TN);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    if (threadIdx.x == 0) {
        bulk_load(pp.wbuf, Wg, CHUNK_BYTES, &pp.bars[0]);
        if (NCH > 1) bulk_load(pp.wbuf + CHUNK_BYTES / 4, Wg + (size_t)KC * N, CHUNK_BYTES, &pp.bars[1]);
    }
#pragma unroll 1
    for (int ch = 0; ch < NCH; ++ch) {
        const int b = ch & 1;
        mbar_wait(&pp.bars[b], (pp.phase >> b) & 1u);
        pp.phase ^= 1u << b;
        const float* wb = pp.wbuf + b * (CHUNK_BYTES / 4);
        const float* a0 = A + (size_t)(ty * TM) * lda + ch * KC;
#pragma unroll 2
        for (int k4 = 0; k4 < KC / 4; ++k4) {
            float4 a[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + (size_t)i * lda + k4 * 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float bv[TN];
                const float* wr = wb + (k4 * 4 + kk) * N + tx * 4;
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    float4 t = *reinterpret_cast<const float4*>(wr + g * GS);
                    bv[g * 4 + 0] = t.x; bv[g * 4 + 1] = t.y; bv[g * 4 + 2] = t.z; bv[g * 4 + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av, bv[j], acc[i][j]);
                }
            }
        }
        __syncthreads();   // every thread is done with buffer b
        if (threadIdx.x == 0 && ch + 2 < NCH)
            bulk_load(pp.wbuf + b * (CHUNK_BYTES / 4), Wg + (size_t)(ch + 2) * KC * N, CHUNK_BYTES, &pp.bars[b]);
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            float4* o = reinterpret_cast<float4*>(OUT + (size_t)(ty * TM + i) * ldo + g * GS + tx * 4);
            float v[4] = {acc[i][g * 4 + 0], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]};
            if (HAS_BIAS) {
                const float4 bb = *reinterpret_cast<const float4*>(bias + g * GS + tx * 4);
                v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
            }
            if (MODE == 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.f);
            }
            if (MODE == 2) {
                const float4 old = *o;
                v[0] = old.x > 0.f ? v[0] : 0.f; v[1] = old.y > 0.f ? v[1] : 0.f;
                v[2] = old.z > 0.f ? v[2] : 0.f; v[3] = old.w > 0.f ? v[3] : 0.f;
            }
            *o = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    __syncthreads();
}

}  // namespace mlp

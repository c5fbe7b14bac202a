// tc_selftest.cu -- one-tile tcgen05 GEMM used by tests/test_tc_gpu.py to pin the descriptor / TMEM conventions
// of tc_tile.cuh (interleaved no-swizzle images, K-major and MN-major operands, M = 64 and M = 128 accumulators).
#include "tc_tile.cuh"

namespace {
using namespace tc;

struct TcTestParams { const float* a; const float* b; float* d; int M, N, K, a_mn, b_mn, a_floats, b_floats; int b_lbo, b_sbo, b_kstep; };

__global__ void __launch_bounds__(128) k_tc_gemm_test(const TcTestParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sa = reinterpret_cast<float*>(smem_raw);
    float* sb = sa + P.a_floats;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    for (int i = threadIdx.x; i < P.a_floats; i += blockDim.x) sa[i] = P.a[i];
    for (int i = threadIdx.x; i < P.b_floats; i += blockDim.x) sb[i] = P.b[i];
    if (threadIdx.x == 0) { mlp::mbar_init(&bar, 1); mlp::fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 256);
    mlp::fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(P.M, P.N, P.a_mn, P.b_mn);
        for (int ks = 0; ks < P.K / 8; ++ks) {
            const uint64_t ad = P.a_mn ? desc_mnmajor(smem_u32(sa) + ks * (P.M * 32), P.M) : desc_kmajor(smem_u32(sa) + ks * 256, P.K);
            uint64_t bd = P.b_mn ? desc_mnmajor(smem_u32(sb) + ks * (P.N * 32), P.N) : desc_kmajor(smem_u32(sb) + ks * 256, P.K);
            if (P.b_lbo) bd = make_desc(smem_u32(sb) + ks * (P.b_kstep & 0xFFFFF), P.b_lbo, P.b_sbo) | ((uint64_t)(P.b_kstep >> 20) << 61);
            mma_tf32(tmem, ad, bd, idesc, ks > 0);
        }
        mma_commit(&bar);
    }
    mlp::mbar_wait(&bar, 0);
    fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < P.N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_wait_ld();
        int row = -1;
        if (P.M == 128) row = warp * 32 + lane;
        else if (lane < 16) row = warp * 16 + lane;          // M = 64: rows 16j+i live in lane 32j+i
        if (row >= 0)
            for (int j = 0; j < 16; ++j) P.d[(size_t)row * P.N + c0 + j] = v[j];
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}
}  // namespace

extern "C" int rl_tc_gemm_test_ex(const float* a_img, const float* b_img, float* d, int M, int N, int K, int a_mn, int b_mn,
                               int b_lbo, int b_sbo, int b_kstep, void* stream) {
    RL_ARG_CHECK(a_img && b_img && d && (M == 64 || M == 128) && N % 16 == 0 && N <= 256 && K % 8 == 0);
    TcTestParams P{a_img, b_img, d, M, N, K, a_mn, b_mn, M * K, N * K, b_lbo, b_sbo, b_kstep};
    const size_t smem = sizeof(float) * (size_t)(P.a_floats + P.b_floats);
    RL_ARG_CHECK(smem <= 200 * 1024);
    RL_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gemm_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tc_gemm_test<<<1, 128, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

extern "C" int rl_tc_gemm_test(const float* a_img, const float* b_img, float* d, int M, int N, int K, int a_mn, int b_mn,
                               void* stream) {
    return rl_tc_gemm_test_ex(a_img, b_img, d, M, N, K, a_mn, b_mn, 0, 0, 0, stream);
}

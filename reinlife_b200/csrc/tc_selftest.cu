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

// ---- fp16 one-tile GEMM with caller-supplied descriptor geometry (pins the K-major / MN-major conventions of kind::f16) ----
namespace {
using namespace tc;
struct TcTestParamsH { const uint16_t* a; const uint16_t* b; float* d; int M, N, K, a_halves, b_halves;
                       uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep; int a_mn, b_mn; uint32_t a_kblk, b_kblk; int a_layout, b_layout; };

__global__ void __launch_bounds__(128) k_tc_gemm_test_h(const TcTestParamsH P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint16_t* sa = reinterpret_cast<uint16_t*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    uint16_t* sb = sa + ((P.a_halves + 511) & ~511);        // 1024-byte aligned (SWIZZLE_128B atoms)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    for (int i = threadIdx.x; i < P.a_halves; i += blockDim.x) sa[i] = P.a[i];
    for (int i = threadIdx.x; i < P.b_halves; i += blockDim.x) sb[i] = P.b[i];
    if (threadIdx.x == 0) { mlp::mbar_init(&bar, 1); mlp::fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 256);
    mlp::fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)P.a_mn << 15) | ((uint32_t)P.b_mn << 16) | ((uint32_t)(P.N >> 3) << 17) | ((uint32_t)(P.M >> 4) << 24);
        for (int ks = 0; ks < P.K / 16; ++ks) {
            // k-step address: linear, or (a_kblk != 0) four steps inside a swizzle atom, then the next K block
            const uint32_t ao = P.a_kblk ? (ks >> 2) * P.a_kblk + (ks & 3) * P.a_kstep : ks * P.a_kstep;
            const uint32_t bo = P.b_kblk ? (ks >> 2) * P.b_kblk + (ks & 3) * P.b_kstep : ks * P.b_kstep;
            const uint64_t ad = make_desc(smem_u32(sa) + ao, P.a_lbo, P.a_sbo) | ((uint64_t)P.a_layout << 61);
            const uint64_t bd = make_desc(smem_u32(sb) + bo, P.b_lbo, P.b_sbo) | ((uint64_t)P.b_layout << 61);
            const uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
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
        else if (lane < 16) row = warp * 16 + lane;
        if (row >= 0)
            for (int j = 0; j < 16; ++j) P.d[(size_t)row * P.N + c0 + j] = v[j];
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}
}  // namespace

extern "C" int rl_tc_gemm_test_hx(const void* a_img, const void* b_img, float* d, int M, int N, int K, int a_halves, int b_halves,
                                  uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                                  int a_mn, int b_mn, uint32_t a_kblk, uint32_t b_kblk, int a_layout, int b_layout, void* stream);
extern "C" int rl_tc_gemm_test_h(const void* a_img, const void* b_img, float* d, int M, int N, int K, int a_halves, int b_halves,
                                 uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                                 int a_mn, int b_mn, void* stream) {
    return rl_tc_gemm_test_hx(a_img, b_img, d, M, N, K, a_halves, b_halves, a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, a_mn, b_mn, 0, 0, 0, 0, stream);
}
extern "C" int rl_tc_gemm_test_hx(const void* a_img, const void* b_img, float* d, int M, int N, int K, int a_halves, int b_halves,
                                  uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                                  int a_mn, int b_mn, uint32_t a_kblk, uint32_t b_kblk, int a_layout, int b_layout, void* stream) {
    RL_ARG_CHECK(a_img && b_img && d && (M == 64 || M == 128) && N % 16 == 0 && N <= 256 && K % 16 == 0);
    TcTestParamsH P{reinterpret_cast<const uint16_t*>(a_img), reinterpret_cast<const uint16_t*>(b_img), d, M, N, K, a_halves, b_halves,
                    a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, a_mn, b_mn, a_kblk, b_kblk, a_layout, b_layout};
    const size_t smem = 2 * (size_t)(((a_halves + 511) & ~511) + b_halves) + 2048;
    RL_ARG_CHECK(smem <= 200 * 1024);
    RL_CUDA_CHECK(cudaFuncSetAttribute(k_tc_gemm_test_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tc_gemm_test_h<<<1, 128, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

// ---- timing probe: cycles per tcgen05.mma kind::tf32 for a given operand layout (data is irrelevant: zeros) ----
namespace {
using namespace tc;
struct TcBenchParams { int M, N, ksteps, mode, iters; long long* out; };

__global__ void __launch_bounds__(128) k_tc_mma_bench(const TcBenchParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    float* sm = reinterpret_cast<float*>(smem_raw);
    for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) sm[i] = 0.f;
    if (threadIdx.x == 0) { mlp::mbar_init(&bar, 1); mlp::fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 256);
    mlp::fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x < 32) {
        // the whole warp runs the (warp-uniform) issue loop so that descriptors live in uniform registers; one elected
        // lane issues the MMAs (P.mode >= 10: elected-lane variant; < 10: `if (threadIdx.x == 0)` variant for comparison)
        const bool uniform_flow = P.mode >= 10 && P.mode < 20;
        const int mode = P.mode % 10;
        uint32_t elected = 0;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
        const bool me = uniform_flow ? (elected != 0) : (threadIdx.x == 0);
        if (uniform_flow || threadIdx.x == 0) {
            const uint32_t sa = smem_u32(sm), sb = smem_u32(sm + 24 * 1024);
            const uint32_t idesc = make_idesc(P.M, P.N, 0, 0);
            uint32_t ph = 0;
            long long t_issue = 0, t_total = 0;
            for (int it = 0; it < P.iters; ++it) {
                const long long t0 = clock64();
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    if (ks < P.ksteps) {
                        uint64_t ad, bd;
                        if (mode == 0) {            // no swizzle, K = 64 floats per row: LBO 128 B, SBO K*32 B
                            ad = desc_kmajor(sa + ks * 256, 64); bd = desc_kmajor(sb + ks * 256, 64);
                        } else if (mode == 1) {     // no swizzle, padded chunk stride (LBO 144 B)
                            ad = make_desc(sa + ks * 288, 144, 16 * 144); bd = make_desc(sb + ks * 288, 144, 16 * 144);
                        } else {                    // SWIZZLE_128B K-major
                            const uint32_t ao = (ks >> 2) * (P.M * 128) + (ks & 3) * 32, bo = (ks >> 2) * (P.N * 128) + (ks & 3) * 32;
                            ad = make_desc(sa + ao, 16, 1024) | (2ull << 61); bd = make_desc(sb + bo, 16, 1024) | (2ull << 61);
                        }
                        if (me) {
                            if (P.mode >= 20) {   // kind::f16 (fp16 operands, K = 16 per instruction = the same 32 bytes), fp32 accumulate
                                const uint32_t idh = (1u << 4) | ((uint32_t)(P.N >> 3) << 17) | ((uint32_t)(P.M >> 4) << 24);
                                const uint32_t acc = ks > 0;
                                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(idh), "r"(acc) : "memory");
                            } else {
                                mma_tf32(tmem, ad, bd, idesc, ks > 0);
                            }
                        }
                    }
                }
                if (me) mma_commit(&bar);
                const long long t1 = clock64();
                mlp::mbar_wait(&bar, ph & 1); ++ph;
                fence_after();
                const long long t2 = clock64();
                if (it > 0) { t_issue += t1 - t0; t_total += t2 - t0; }
            }
            if (me) { P.out[0] = t_issue / (P.iters - 1); P.out[1] = t_total / (P.iters - 1); }
        }
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}
}  // namespace

namespace {
// n_issuers warps (lane 0 of each) issue `ksteps` MMAs each, concurrently, into disjoint TMEM column ranges
__global__ void __launch_bounds__(128) k_tc_mma_bench_par(const TcBenchParams P, int n_issuers) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ long long t_end[4];
    float* sm = reinterpret_cast<float*>(smem_raw);
    for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) sm[i] = 0.f;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mlp::mbar_init(&bar[i], 1); mlp::fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 512);
    mlp::fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    const int warp = threadIdx.x >> 5;
    long long acc = 0;
    for (int it = 0; it < P.iters; ++it) {
        __syncthreads();
        const long long t0 = clock64();
        if ((threadIdx.x & 31) == 0 && warp < n_issuers) {
            const uint32_t sa = smem_u32(sm), sb = smem_u32(sm + 24 * 1024);
            const uint32_t idesc = make_idesc(P.M, P.N, 0, 0);
            for (int ks = 0; ks < P.ksteps; ++ks)
                mma_tf32(tmem + warp * 128, desc_kmajor(sa + ks * 256, 64), desc_kmajor(sb + ks * 256, 64), idesc, ks > 0);
            mma_commit(&bar[warp]);
            mlp::mbar_wait(&bar[warp], it & 1);
            fence_after();
            t_end[warp] = clock64();
        }
        __syncthreads();
        if (threadIdx.x == 0 && it > 0) {
            long long m = 0;
            for (int w = 0; w < n_issuers; ++w) m = t_end[w] > m ? t_end[w] : m;
            acc += m - t0;
        }
    }
    if (threadIdx.x == 0) { P.out[0] = acc / (P.iters - 1); P.out[1] = acc / (P.iters - 1); }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
}  // namespace

// out_host[0] = average cycles to issue `ksteps` MMAs + commit, out_host[1] = average cycles until they completed
extern "C" int rl_tc_mma_bench(int M, int N, int ksteps, int mode, int iters, long long* out_host) {
    RL_ARG_CHECK(out_host && (M == 64 || M == 128) && N % 16 == 0 && N <= 256 && ksteps > 0 && ksteps <= 8 && iters > 1 && mode >= 0 && mode < 34);
    long long* dev = nullptr;
    RL_CUDA_CHECK(cudaMalloc(&dev, 2 * sizeof(long long)));
    TcBenchParams P{M, N, ksteps, mode, iters, dev};
    const size_t smem = 192 * 1024;
    RL_CUDA_CHECK(cudaFuncSetAttribute(k_tc_mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (mode >= 30) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_tc_mma_bench_par, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_tc_mma_bench_par<<<1, 128, smem, 0>>>(P, mode - 30 + 1);
    } else
    k_tc_mma_bench<<<1, 128, smem, 0>>>(P);
    RL_CUDA_CHECK(cudaDeviceSynchronize());
    RL_CUDA_CHECK(cudaMemcpy(out_host, dev, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(dev);
    return RL_OK;
}

// tc_issue_probe.cu -- timing probe: how fast can ONE warp feed tcgen05.mma kind::f16 (M = 128) to the tensor pipe?
// Two issue styles over the same straight-line MMA sequence (data is irrelevant: zeros):
//   style 0: `if (lane == 0)` -- the issuing thread is picked by a per-thread predicate, descriptors live in vector
//            registers (ptxas wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY convergence loop);
//   style 1: the whole warp runs the loop in warp-uniform control flow, descriptors are built from uniform values only
//            and the MMA is guarded by elect.sync (descriptors stay in uniform registers, UTCHMMA issues directly).
// layout 0: K-major no-swizzle (LBO 128 B, SBO = width * 16 B), 1: MN-major no-swizzle, 2: K-major SWIZZLE_128B.
// scripts/tc_issue_probe.py prints cycles per MMA (issue side and completion side).
#include "tc_tile.cuh"

namespace {
using namespace tc;

struct IssueProbeParams { int N, nmma, layout, style, iters, n_warps; long long* out; };

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128) k_tc_issue_probe(const IssueProbeParams P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ long long t_out[4][3];
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem_raw);
    for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) sm[i] = 0u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mlp::mbar_init(&bar[i], 1); mlp::fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 512);
    mlp::fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem_raw);
    // operand descriptors of k-step 0 and the per-k-step advance (in 16-byte descriptor units)
    uint64_t a0, b0; uint32_t astep, bstep;
    const uint32_t sa = sbase, sb = sbase + 64 * 1024;
    if (P.layout == 0) {            // images of width 128 halves: LBO 128 B, SBO 2048 B, k-step 256 B
        a0 = make_desc(sa, 128u, 2048u); b0 = make_desc(sb, 128u, 2048u); astep = 16u; bstep = 16u;
    } else if (P.layout == 1) {     // MN-major: A image [k][128], B image [k][N]: SBO 128 B, LBO = W * 16 B, k-step = 2 W * 16 B
        a0 = make_desc(sa, 128u * 16u, 128u); b0 = make_desc(sb, (uint32_t)P.N * 16u, 128u); astep = 2u * 128u; bstep = 2u * (uint32_t)P.N;
    } else {                        // SWIZZLE_128B K-major atoms of 8 rows x 128 B: SBO 1024 B, k-step 32 B inside the atom
        a0 = make_desc(sa, 16u, 1024u) | (2ull << 61); b0 = make_desc(sb, 16u, 1024u) | (2ull << 61); astep = 2u; bstep = 2u;
    }
    const uint32_t idesc = (1u << 4) | ((uint32_t)(P.layout == 1) << 15) | ((uint32_t)(P.layout == 1) << 16) |
                           ((uint32_t)(P.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t dcol = tmem + (uint32_t)warp * (P.n_warps > 1 ? 256u / P.n_warps * 2u : 0u);
    long long t_issue = 0, t_total = 0;
    if (warp < P.n_warps) {
        if (P.style == 0) {
            if (lane == 0) {
                for (int it = 0; it < P.iters; ++it) {
                    const long long t0 = clock64();
                    uint64_t a = a0, b = b0;
#pragma unroll 1
                    for (int i = 0; i < P.nmma; ++i) {
                        mma_f16(dcol, a, b, idesc, i != 0);
                        a += astep; b += bstep;
                        if ((i & 3) == 3) { a = a0; b = b0; }
                    }
                    mma_commit(&bar[warp]);
                    const long long t1 = clock64();
                    mlp::mbar_wait(&bar[warp], it & 1);
                    fence_after();
                    const long long t2 = clock64();
                    if (it > 0) { t_issue += t1 - t0; t_total += t2 - t0; }
                }
                t_out[warp][0] = t_issue; t_out[warp][1] = t_total;
            }
        } else {
            for (int it = 0; it < P.iters; ++it) {
                const long long t0 = clock64();
                const bool me = elect_one();
                uint64_t a = a0, b = b0;
#pragma unroll 1
                for (int i = 0; i < P.nmma; i += 4) {
                    if (me) {
                        mma_f16(dcol, a, b, idesc, i != 0);
                        mma_f16(dcol, a + astep, b + bstep, idesc, 1u);
                        mma_f16(dcol, a + 2 * astep, b + 2 * bstep, idesc, 1u);
                        mma_f16(dcol, a + 3 * astep, b + 3 * bstep, idesc, 1u);
                    }
                }
                if (me) mma_commit(&bar[warp]);
                __syncwarp();
                const long long t1 = clock64();
                mlp::mbar_wait(&bar[warp], it & 1);
                fence_after();
                const long long t2 = clock64();
                if (it > 0) { t_issue += t1 - t0; t_total += t2 - t0; }
            }
            if (lane == 0) { t_out[warp][0] = t_issue; t_out[warp][1] = t_total; }
        }
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        long long mi = 0, mt = 0;
        for (int w = 0; w < P.n_warps; ++w) { mi = t_out[w][0] > mi ? t_out[w][0] : mi; mt = t_out[w][1] > mt ? t_out[w][1] : mt; }
        P.out[0] = mi / (P.iters - 1); P.out[1] = mt / (P.iters - 1);
    }
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}
}  // namespace

// out_host[0] = average cycles to issue `nmma` MMAs + commit, out_host[1] = average cycles until they completed
extern "C" int rl_tc_issue_probe(int N, int nmma, int layout, int style, int n_warps, int iters, long long* out_host) {
    RL_ARG_CHECK(out_host && N % 16 == 0 && N >= 16 && N <= 256 && nmma > 0 && nmma % 4 == 0 && iters > 1 && layout >= 0 && layout <= 2 &&
                 (style == 0 || style == 1) && n_warps >= 1 && n_warps <= 4);
    long long* dev = nullptr;
    RL_CUDA_CHECK(cudaMalloc(&dev, 2 * sizeof(long long)));
    IssueProbeParams P{N, nmma, layout, style, iters, n_warps, dev};
    const size_t smem = 192 * 1024;
    RL_CUDA_CHECK(cudaFuncSetAttribute(k_tc_issue_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tc_issue_probe<<<1, 128, smem, 0>>>(P);
    RL_CUDA_CHECK(cudaDeviceSynchronize());
    RL_CUDA_CHECK(cudaMemcpy(out_host, dev, 2 * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(dev);
    return RL_OK;
}

// ---- TMA tile::gather4 probe: 128 ring rows (fp16, 160 halves) by index -> three [128 rows][128 B] SWIZZLE_128B K blocks ----
#include "tma_util.cuh"
namespace {
__global__ void __launch_bounds__(128) k_tma_gather_test(const __grid_constant__ CUtensorMap map, const int32_t* __restrict__ idx, uint4* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    unsigned char* img = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 3 * 16384 / 16; i += blockDim.x) reinterpret_cast<uint4*>(img)[i] = make_uint4(0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu);
    if (threadIdx.x == 0) { mlp::mbar_init(&bar, 1); mlp::fence_mbar_init(); }
    mlp::fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(3 * 16384) : "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        const int g = threadIdx.x;
        const int r0 = idx[4 * g], r1 = idx[4 * g + 1], r2 = idx[4 * g + 2], r3 = idx[4 * g + 3];
        for (int kb = 0; kb < 3; ++kb) tma::gather4(smem_u32(img) + kb * 16384 + g * 512, &map, smem_u32(&bar), kb * 64, r0, r1, r2, r3);
    }
    mlp::mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 3 * 16384 / 16; i += blockDim.x) out[i] = reinterpret_cast<const uint4*>(img)[i];
}
}  // namespace

// out = the 48 KB shared-memory image after gathering rows idx[0..127] of the fp16 tensor [n_rows][160] at `ring`
extern "C" int rl_tma_gather_test(const void* ring, long long n_rows, const int32_t* idx, void* out, int box_rows) {
    RL_ARG_CHECK(ring && idx && out && n_rows > 0 && box_rows >= 1);
    CUtensorMap map;
    const int rc = tma::make_rows_map(&map, ring, (uint64_t)n_rows, 160, (uint32_t)box_rows);
    if (rc != 0) return rl_set_err(RL_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    const size_t smem = 3 * 16384 + 2048;
    RL_CUDA_CHECK(cudaFuncSetAttribute(k_tma_gather_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tma_gather_test<<<1, 128, smem, 0>>>(map, idx, reinterpret_cast<uint4*>(out));
    RL_CUDA_CHECK(cudaDeviceSynchronize());
    return RL_OK;
}

// tc_bm.cuh -- building blocks of the BATCH-MAJOR tcgen05 kernels (tc_pair_kernels.cu: train() events, tc_act_kernels.cu: get_action):
// a tile is M = 128 batch rows, D[batch][features] = Act[128][K] * W[features][K]^T with fp16 operands (kind::f16) and fp32
// accumulators in TMEM.  Operand images: activations written by the epilogue live in the interleaved no-swizzle layout (8 rows x 16 B
// core matrices; the same image reads K-major and MN-major); input rows gathered from global memory live in SWIZZLE_128B K blocks
// ([128 rows][128 B], 16-byte unit ^= row & 7 -- what a TMA tile::gather4 writes), read K-major (A of L1) or MN-major (B of dW1).
// scripts/tc_probe_h.py, scripts/tc_probe_sw128.py and scripts/tc_issue_probe.py pin these conventions on the hardware.
#pragma once
#include <cuda_fp16.h>
#include "tc_tile.cuh"

namespace bm {

using namespace tc;

constexpr int PB = 128;                  // rows per tile
constexpr int HCH = 4096;                // halves per weight chunk (8 KB)
constexpr int XBLK = PB * 128;           // bytes of one SWIZZLE_128B K block: 128 rows x 64 halves
constexpr int XIMG = 3 * XBLK;           // 160 halves = 2.5 blocks; the upper half of block 2 is never read

// interleaved no-swizzle fp16 image of width K: 8 rows x 16 bytes core matrices
__device__ __forceinline__ int himg(int r, int c, int K) { return (r >> 3) * (K * 8) + (c >> 3) * 64 + (r & 7) * 8 + (c & 7); }
// byte offset of 16-byte unit `u` (8 halves) of row r in a SWIZZLE_128B image of [128][64-half] K blocks
__device__ __forceinline__ int ximg(int r, int u) { return (u >> 3) * XBLK + r * 128 + (((u & 7) ^ (r & 7)) << 4); }
// K-major no-swizzle reading (rows = M / N index): LBO = next 8 k (128 B), SBO = next 8 rows (K * 16 B); one K = 16 step = +256 B
__device__ __forceinline__ uint64_t dk(uint32_t addr, int K) { return make_desc(addr, 128u, (uint32_t)K * 16u); }
// MN-major no-swizzle reading (rows = k index, columns = M / N index): SBO = next 8 columns (128 B), LBO = next 8 rows (W * 16 B);
// one K = 16 step = +2 * W * 16 B
__device__ __forceinline__ uint64_t dm(uint32_t addr, int W) { return make_desc(addr, (uint32_t)W * 16u, 128u); }
// SWIZZLE_128B image, K-major reading (A of L1): 8-row groups 1024 B apart; a K = 16 step is +32 B inside the 128-byte row,
// the next 64 halves are the next K block
__device__ __forceinline__ uint64_t dxk(uint32_t addr, int ks) {
    return make_desc(addr + (uint32_t)(ks >> 2) * XBLK + (uint32_t)(ks & 3) * 32u, 16u, 1024u) | (2ull << 61);
}
// SWIZZLE_128B image, MN-major reading (B of dW1: N = image columns, K = image rows): LBO = next 64 columns (one K block),
// SBO = next 8 rows; a K = 16 step is +2048 B
__device__ __forceinline__ uint64_t dxm(uint32_t addr, int ks) { return make_desc(addr + (uint32_t)ks * 2048u, (uint32_t)XBLK, 1024u) | (2ull << 61); }
__host__ __device__ constexpr uint32_t idesc_h(int M, int N, int a_mn, int b_mn) {      // kind::f16: fp16 A/B, fp32 D
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_h(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t pk(float a, float b) {      // one F2FP.PACK_AB (full rate) instead of two F2F (quarter rate, XU pipe)
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));     // upper half <- b, lower half <- a
    return r;
}
// relu(a), relu(b) -> packed halves (F2FP.RELU); finite-saturating pack for the scaled backward operands (F2FP.SATFINITE)
__device__ __forceinline__ uint32_t pk_relu(float a, float b) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ uint32_t pk_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// per half: v where h > 0, else 0 (ReLU derivative applied to a packed pair: HSET2.BF.GT + HMUL2)
__device__ __forceinline__ uint32_t mask_pos(uint32_t v, uint32_t h) {
    const __half2 m = __hgt2(*reinterpret_cast<const __half2*>(&h), __float2half2_rn(0.f));
    const __half2 r = __hmul2(*reinterpret_cast<const __half2*>(&v), m);
    return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {      // (barrier armed by the caller)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void head_bar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }


}  // namespace bm

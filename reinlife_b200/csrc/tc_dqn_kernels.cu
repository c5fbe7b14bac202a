// tc_dqn_kernels.cu -- k_learn_dqn_p<MODE>: one iteration of train(q, q_target, memory, optimizer) (Models/DQN.py:142-153; MODE 1:
// PERDQN's train_model, Models/PERDQN.py:130-186) for every train event of the brain, on the 5th-gen tensor cores in the
// batch-major form of tc_bm.cuh.  Same contract and outputs as k_learn_dqn<MODE> (learn_rows_kernels.cu), fp16 operands / fp32
// accumulation (the tolerance of the dueling tensor-core kernels, tests/test_scale_gpu.py).
//
// A tile is 128 sampled rows (four 32-row DQN events / two 64-row PERDQN events; rows never interact).  The 153-128-64-8 network is
// small enough that BOTH nets' fp16 operand images stay in shared memory for the whole launch (2 x 58 KB, converted from the
// float32 parameters once per CTA: no weight stream, no image kernels) and ALL weight gradients stay in TMEM until the CTA is done
// (dW1 160 + dW2 64 + dWh 16 + db2 16 columns: no reds):
//   target: L1 X'.W1t^T -> H1, L2 -> H2, head -> max_a q_target(s')          X', X: SWIZZLE_128B images gathered float32 -> fp16 by
//   eval:   L1 X.W1^T  -> H1, L2 -> H2, head -> q(s)[a], loss, dOut           eight gather warps (the rings of these brains are float32)
//   back:   dH2 = dOut.Wh (in place of H2, masked), dWh += H2^T dOut (M = 64), dW2 += H1^T dH2, dH1 = dH2.W2 (masked),
//           db2 += ones^T dH2 (M = 64), dW1 += dH1^T X (column 159 of X is 1: db1), dbh by shuffles (8 columns)
// One warp-uniform MMA issuer (elect.sync); stages handed over with go / done mbarriers, never two completions of one barrier
// without a wait in between.
#include <string.h>
#include "tc_bm.cuh"
#include "models.cuh"

int rl_count_valid_launch(const int32_t* sample_idx, int batch, const int32_t* ev_total, float* out, void* stream);   // learn_rows_kernels.cu

namespace {

using namespace tc;
using namespace bm;
using mlp::mbar_init; using mlp::mbar_wait; using mlp::fence_mbar_init; using mlp::fence_proxy_async;

constexpr int NEPI = 256;                // epilogue threads (warps 0-7)
constexpr int NGA = 8;                   // gather warps 10..17
constexpr int NTH = NEPI + 64 + 32 * NGA;    // + warp 8 (TMEM allocation) + MMA issuer (warp 9)
constexpr int N1 = 128, N2 = 64;
constexpr float H_SCALE = 256.0f;        // backward operands are scaled by 2^8 (exact), removed when gradients leave TMEM

// ---- shared memory (bytes, from a 1024-byte aligned base) ----
constexpr int DO_X = 0;                                   // X' / X, SWIZZLE_128B
constexpr int DO_H1 = DO_X + XIMG;                        // H1 -> dH1 [128][128]
constexpr int DO_H2 = DO_H1 + PB * N1 * 2;                // H2 -> dH2 [128][64]
constexpr int DO_DOUT = DO_H2 + PB * N2 * 2;              // dOut [128][16]
constexpr int DO_ONES = DO_DOUT + PB * 16 * 2;            // [16 k][16] halves of 1.0
constexpr int WNET = (N1 * 160 + N2 * N1 + 16 * N2) * 2;  // one net: W1 [128][160], W2 [64][128], Wh [16][64] halves
constexpr int DO_W = DO_ONES + 512;                       // [2] nets: 0 target, 1 eval
constexpr int DO_BIAS = DO_W + 2 * WNET;                  // [2] x { b1[128] b2[64] bh[16] } floats
constexpr int DO_BARS = DO_BIAS + 2 * 4 * 208;
constexpr int NBAR = 8;                                   // done doneL1 go xfullA xfullB xfreeA xfreeB (+1 spare)
constexpr size_t DQP_SMEM = DO_BARS + 8 * NBAR + 16 + 1024;
static_assert(DQP_SMEM <= 227 * 1024 && DO_BARS % 8 == 0 && DO_W % 128 == 0, "shared memory budget");
constexpr int WO_W2 = N1 * 160 * 2, WO_WH = WO_W2 + N2 * N1 * 2;      // byte offsets inside a net's block

// TMEM columns: resident gradients, then the work area
constexpr int TC_DW1 = 0, TC_DW2 = 160, TC_DWH = 224, TC_DB2 = 240, TC_L1 = 256, TC_L2 = 384, TC_HD = 448;

struct DqnParams {
    rl_world_cfg cfg;
    const int32_t* ev_rows;
    const int32_t* ev_total;
    rl_replay_bufs rp;
    const int32_t* sample_idx;
    const float* ev_weight;
    rl_learn_bufs lb;
};

template <int MODE>
__global__ void __launch_bounds__(NTH, 1) k_learn_dqn_p(const DqnParams P) {
    constexpr int EB = MODE == 0 ? 32 : 64;              // rows per event
    using L = Layout<RL_MODEL_DQN>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __half* sH1 = reinterpret_cast<__half*>(smem + DO_H1);
    __half* sH2 = reinterpret_cast<__half*>(smem + DO_H2);
    __half* sDout = reinterpret_cast<__half*>(smem + DO_DOUT);
    float* bias = reinterpret_cast<float*>(smem + DO_BIAS);            // [net][208]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DO_BARS);
    uint64_t* done = bars; uint64_t* doneL1 = bars + 1; uint64_t* go = bars + 2;
    // xfullA / xfullB: the target / eval rows of a tile are in the X image; xfreeA: dW1 of the previous tile has read X (the next
    // tile's target rows may overwrite it); xfreeB: the target L1 has read X' (this tile's eval rows may overwrite it)
    uint64_t* xfullA = bars + 3; uint64_t* xfullB = bars + 4; uint64_t* xfreeA = bars + 5; uint64_t* xfreeB = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* G = P.lb.grad_scratch + (size_t)blockIdx.x * L::N_TRAIN;
    const int total = *P.ev_total;
    const int n_tiles = (total * EB + PB - 1) / PB;
    const int n_my = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        mbar_init(done, 1); mbar_init(doneL1, 1); mbar_init(go, NEPI);
        mbar_init(xfullA, 32 * NGA); mbar_init(xfullB, 32 * NGA); mbar_init(xfreeA, 1); mbar_init(xfreeB, 1);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    // operand images of both nets (interleaved no-swizzle, K-major: W1 [n = k1][k = x], W2 [n = n2][k = k1], Wh [n = j][k = n2]),
    // biases, the ones block, the zeroed gradient slab
    for (int net = 0; net < 2; ++net) {
        const float* p = net ? P.lb.params : P.lb.target;
        __half* w1 = reinterpret_cast<__half*>(smem + DO_W + net * WNET);
        __half* w2 = reinterpret_cast<__half*>(smem + DO_W + net * WNET + WO_W2);
        __half* wh = reinterpret_cast<__half*>(smem + DO_W + net * WNET + WO_WH);
        for (int i = threadIdx.x; i < 160 * N1; i += NTH) { const int k = i / N1, n = i - k * N1; w1[himg(n, k, 160)] = __float2half_rn(p[L::OFF_W1T + i]); }
        for (int i = threadIdx.x; i < N1 * N2; i += NTH) { const int k = i / N2, n = i - k * N2; w2[himg(n, k, N1)] = __float2half_rn(p[L::OFF_W2T + i]); }
        for (int i = threadIdx.x; i < N2 * 16; i += NTH) {
            const int k = i >> 4, j = i & 15;
            wh[himg(j, k, N2)] = __float2half_rn(j < 8 ? p[L::OFF_WH + k * 8 + j] : 0.f);
        }
        for (int i = threadIdx.x; i < 208; i += NTH)
            bias[net * 208 + i] = i < 128 ? p[L::OFF_B1 + i] : i < 192 ? p[L::OFF_B2 + (i - 128)] : i < 200 ? p[L::OFF_BH + (i - 192)] : 0.f;
    }
    if (threadIdx.x < 256) reinterpret_cast<__half*>(smem + DO_ONES)[threadIdx.x] = __float2half_rn(1.0f);
    for (int i = threadIdx.x; i < L::N_TRAIN / 4; i += NTH) reinterpret_cast<float4*>(G)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t aX = smem_u32(smem + DO_X), aH1 = smem_u32(sH1), aH2 = smem_u32(sH2), aD = smem_u32(sDout), aOnes = smem_u32(smem + DO_ONES);
    const uint32_t aW = smem_u32(smem + DO_W);
    const int S = P.cfg.slot_cap, cap = P.rp.capacity;

    if (warp == 9) {
        // =================================== MMA issuer (one warp, warp-uniform) ===================================
        const bool me = elect_one();
        const uint32_t T0 = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        const int n_u = __shfl_sync(0xffffffffu, n_my, 0);
        uint32_t go_no = 0;
        auto wait_go = [&]() { mbar_wait(go, go_no & 1); ++go_no; fence_after(); };
        auto commit = [&](uint64_t* bar) { if (me) mma_commit(bar); };
        // forward of one net: L1 (10 k-steps, N = 128), L2 (8 k-steps, N = 64), head (4 k-steps, N = 16), each behind a go
        auto l1 = [&](uint32_t aw) {
            const uint32_t id = idesc_h(128, 128, 0, 0);
            uint64_t b = dk(aw, 160);
#pragma unroll 1
            for (int ks = 0; ks < 10; ks += 2) {
                if (me) { mma_h(T0 + TC_L1, dxk(aX, ks), b, id, ks != 0); mma_h(T0 + TC_L1, dxk(aX, ks + 1), b + 16u, id, 1u); }
                b += 32u;
            }
            commit(doneL1);
        };
        auto l2 = [&](uint32_t aw) {
            const uint32_t id = idesc_h(128, N2, 0, 0);
            uint64_t a = dk(aH1, N1), b = dk(aw + WO_W2, N1);
#pragma unroll 1
            for (int ks = 0; ks < 8; ks += 4) {
                if (me) {
                    mma_h(T0 + TC_L2, a, b, id, ks != 0); mma_h(T0 + TC_L2, a + 16u, b + 16u, id, 1u);
                    mma_h(T0 + TC_L2, a + 32u, b + 32u, id, 1u); mma_h(T0 + TC_L2, a + 48u, b + 48u, id, 1u);
                }
                a += 64u; b += 64u;
            }
            commit(done);
        };
        auto head = [&](uint32_t aw) {
            const uint32_t id = idesc_h(128, 16, 0, 0);
            const uint64_t a = dk(aH2, N2), b = dk(aw + WO_WH, N2);
            if (me) {
                mma_h(T0 + TC_HD, a, b, id, 0u); mma_h(T0 + TC_HD, a + 16u, b + 16u, id, 1u);
                mma_h(T0 + TC_HD, a + 32u, b + 32u, id, 1u); mma_h(T0 + TC_HD, a + 48u, b + 48u, id, 1u);
            }
            commit(done);
        };
        for (int t = 0; t < n_u; ++t) {
            const uint32_t first = t == 0 ? 0u : 1u;          // accumulate flag of the resident gradients
            mbar_wait(xfullA, t & 1); fence_after();
            l1(aW);                                            // target net
            commit(xfreeB);
            wait_go(); l2(aW);
            wait_go(); head(aW);
            mbar_wait(xfullB, t & 1); fence_after();
            l1(aW + WNET);                                     // eval net (its L1 runs under the target head epilogue)
            wait_go(); l2(aW + WNET);
            wait_go(); head(aW + WNET);
            wait_go();
            {   // dH2[128 b][64 n2] = dOut[128][16] Wh[16 j][64 n2] (Wh image read MN-major) -> L2 columns; dWh[n2][j] += H2^T dOut (M = 64)
                if (me) mma_h(T0 + TC_L2, dk(aD, 16), dm(aW + WNET + WO_WH, N2), idesc_h(128, N2, 0, 1), 0u);
                const uint32_t id = idesc_h(64, 16, 1, 1);
                uint64_t a = dm(aH2, N2), b = dm(aD, 16);
#pragma unroll 1
                for (int ks = 0; ks < 8; ks += 4) {
                    if (me) {
                        mma_h(T0 + TC_DWH, a, b, id, (first | (uint32_t)ks) != 0u); mma_h(T0 + TC_DWH, a + 128u, b + 32u, id, 1u);
                        mma_h(T0 + TC_DWH, a + 256u, b + 64u, id, 1u); mma_h(T0 + TC_DWH, a + 384u, b + 96u, id, 1u);
                    }
                    a += 512u; b += 128u;
                }
                commit(done);
            }
            wait_go();
            {   // dW2[k1][n2] += H1^T dH2 (both MN-major); dH1[128 b][128 k1] = dH2[128][64] W2[64 n2][128 k1] (W2 image MN-major) -> L1
                // columns; db2[n2] += ones^T dH2 (M = 64, all 16 columns equal)
                const uint32_t id2 = idesc_h(128, N2, 1, 1);
                uint64_t a2 = dm(aH1, N1), b2 = dm(aH2, N2);
#pragma unroll 1
                for (int ks = 0; ks < 8; ks += 4) {
                    if (me) {
                        mma_h(T0 + TC_DW2, a2, b2, id2, (first | (uint32_t)ks) != 0u); mma_h(T0 + TC_DW2, a2 + 256u, b2 + 128u, id2, 1u);
                        mma_h(T0 + TC_DW2, a2 + 512u, b2 + 256u, id2, 1u); mma_h(T0 + TC_DW2, a2 + 768u, b2 + 384u, id2, 1u);
                    }
                    a2 += 1024u; b2 += 512u;
                }
                const uint32_t id = idesc_h(128, 128, 0, 1);
                const uint64_t a = dk(aH2, N2), b = dm(aW + WNET + WO_W2, N1);
                if (me) {
                    mma_h(T0 + TC_L1, a, b, id, 0u); mma_h(T0 + TC_L1, a + 16u, b + 256u, id, 1u);
                    mma_h(T0 + TC_L1, a + 32u, b + 512u, id, 1u); mma_h(T0 + TC_L1, a + 48u, b + 768u, id, 1u);
                }
                const uint32_t idb = idesc_h(64, 16, 1, 1);
                uint64_t ab = dm(aH2, N2);
                const uint64_t bo = dm(aOnes, 16);
#pragma unroll 1
                for (int ks = 0; ks < 8; ks += 4) {
                    if (me) {
                        mma_h(T0 + TC_DB2, ab, bo, idb, (first | (uint32_t)ks) != 0u); mma_h(T0 + TC_DB2, ab + 128u, bo, idb, 1u);
                        mma_h(T0 + TC_DB2, ab + 256u, bo, idb, 1u); mma_h(T0 + TC_DB2, ab + 384u, bo, idb, 1u);
                    }
                    ab += 512u;
                }
                commit(done);
            }
            wait_go();
            {   // dW1[k1][x] += dH1^T X: dH1 (in the H1 region) MN-major, X (SWIZZLE_128B) MN-major, N = 160
                const uint32_t id = idesc_h(128, 160, 1, 1);
                uint64_t a = dm(aH1, N1);
#pragma unroll 1
                for (int ks = 0; ks < 8; ks += 4) {
                    if (me) {
                        mma_h(T0 + TC_DW1, a, dxm(aX, ks), id, (first | (uint32_t)ks) != 0u); mma_h(T0 + TC_DW1, a + 256u, dxm(aX, ks + 1), id, 1u);
                        mma_h(T0 + TC_DW1, a + 512u, dxm(aX, ks + 2), id, 1u); mma_h(T0 + TC_DW1, a + 768u, dxm(aX, ks + 3), id, 1u);
                    }
                    a += 1024u;
                }
                commit(xfreeA);
            }
        }
        commit(done);                                          // everything issued has completed: the resident gradients may be read
    } else if (warp >= 10) {
        // =================================== row gatherers (eight warps) ===================================
        // 128 rows x 20 units of 8 float32 columns -> packed fp16, 10 units per thread; a quarter-warp takes 8 consecutive units of a
        // row.  Column 159 := 1 (db1 rides the dW1 GEMM; the matching W1 rows are structural zeros).  Rows of skipped events
        // (sample_idx < 0) and rows past the last event read transition 0: their weight is 0.
        const int gt = threadIdx.x - 320;
        for (int t = 0; t < n_my; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
#pragma unroll 1
            for (int img = 0; img < 2; ++img) {
                const float* src = img ? P.rp.obs : P.rp.next_obs;
                float4 xa[10], xb[10];
#pragma unroll
                for (int u = 0; u < 10; ++u) {
                    const int v = gt + u * (32 * NGA);
                    const int r = v / 20, oct = v - r * 20;
                    const int gr = tile * PB + r, e = gr / EB;
                    const int si = e < total ? __ldg(P.sample_idx + gr) : -1;
                    const int gi = si >= 0 ? (__ldg(P.ev_rows + e) / S) * cap + si : 0;
                    const float4* g = reinterpret_cast<const float4*>(src + (size_t)gi * RL_K1) + oct * 2;
                    xa[u] = __ldg(g); xb[u] = __ldg(g + 1);
                    if (oct == 19) xb[u].w = 1.0f;
                }
                if (img == 0) { if (t > 0) mbar_wait(xfreeA, (t - 1) & 1); }
                else mbar_wait(xfreeB, t & 1);
#pragma unroll
                for (int u = 0; u < 10; ++u) {
                    const int v = gt + u * (32 * NGA);
                    const int r = v / 20, oct = v - r * 20;
                    *reinterpret_cast<uint4*>(smem + DO_X + ximg(r, oct)) =
                        make_uint4(pk(xa[u].x, xa[u].y), pk(xa[u].z, xa[u].w), pk(xb[u].x, xb[u].y), pk(xb[u].z, xb[u].w));
                }
                fence_proxy_async();
                mbar_arrive(img ? xfullB : xfullA);
            }
        }
    } else if (warp < 8) {
        // =================================== epilogue warps ===================================
        const uint32_t T0 = *tmem_slot;
        uint32_t done_no = 0, l1_no = 0;
        const int q = warp & 3, hh = warp >> 2;
        const int row = q * 32 + lane;                    // batch row of the tile == TMEM lane
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        auto go_signal = [&]() { fence_proxy_async(); fence_before(); mbar_arrive(go); };
        auto wait_done = [&]() { mbar_wait(done, done_no & 1); ++done_no; fence_after(); };
        auto wait_l1 = [&]() { mbar_wait(doneL1, l1_no & 1); ++l1_no; fence_after(); };
        auto relu_store32 = [&](float (&v)[32], const float* b, __half* img, int c0, int K) {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                const float4 b0 = *reinterpret_cast<const float4*>(b + j8 * 8), b1 = *reinterpret_cast<const float4*>(b + j8 * 8 + 4);
                *reinterpret_cast<uint4*>(img + himg(row, c0 + j8 * 8, K)) =
                    make_uint4(pk_relu(v[j8 * 8] + b0.x, v[j8 * 8 + 1] + b0.y), pk_relu(v[j8 * 8 + 2] + b0.z, v[j8 * 8 + 3] + b0.w),
                               pk_relu(v[j8 * 8 + 4] + b1.x, v[j8 * 8 + 5] + b1.y), pk_relu(v[j8 * 8 + 6] + b1.z, v[j8 * 8 + 7] + b1.w));
            }
        };
        auto mask_store32 = [&](float (&v)[32], __half* img, int c0, int K) {      // dH = H > 0 ? acc : 0, in place of H
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                uint4* ph = reinterpret_cast<uint4*>(img + himg(row, c0 + j8 * 8, K));
                const uint4 h4 = *ph;
                *ph = make_uint4(mask_pos(pk_sat(v[j8 * 8], v[j8 * 8 + 1]), h4.x), mask_pos(pk_sat(v[j8 * 8 + 2], v[j8 * 8 + 3]), h4.y),
                                 mask_pos(pk_sat(v[j8 * 8 + 4], v[j8 * 8 + 5]), h4.z), mask_pos(pk_sat(v[j8 * 8 + 6], v[j8 * 8 + 7]), h4.w));
            }
        };
        auto forward_epilogues = [&](const float* b, float (&out)[8]) {   // L1 -> H1, L2 -> H2, head -> out[] (warps 0-3)
            wait_l1();
            {
                const int c0 = hh * 64;
                float v0[32], v1[32];
                tmem_ld32(T0 + t_lane + TC_L1 + c0, v0);
                tmem_ld32(T0 + t_lane + TC_L1 + c0 + 32, v1);
                tmem_wait_ld();
                relu_store32(v0, b + c0, sH1, c0, N1);
                relu_store32(v1, b + c0 + 32, sH1, c0 + 32, N1);
            }
            go_signal();                                            // -> L2
            wait_done();
            {
                const int c0 = hh * 32;
                float v0[32];
                tmem_ld32(T0 + t_lane + TC_L2 + c0, v0);
                tmem_wait_ld();
                relu_store32(v0, b + 128 + c0, sH2, c0, N2);
            }
            go_signal();                                            // -> head
            wait_done();
            if (hh == 0) {
                float v[16];
                tmem_ld16(T0 + t_lane + TC_HD, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) out[j] = v[j] + b[192 + j];
            }
        };
        float acc_bh = 0.f;                                         // lane j < 8 of warps 0-3: sum over the warp's rows of dOut[.][j]
        for (int t = 0; t < n_my; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const int gr = tile * PB + row, e = gr / EB;
            // transition of this row (warps 0-3): action, reward, done mask, weight
            int a_r = 0; float rew = 0.f, dmask = 0.f, wt = 0.f;
            if (hh == 0) {
                const int si = e < total ? P.sample_idx[gr] : -1;
                if (si >= 0) {
                    const size_t gi = (size_t)(P.ev_rows[e] / S) * cap + si;
                    a_r = P.rp.action[gi]; rew = P.rp.reward[gi]; dmask = P.rp.done[gi] ? 0.f : 1.f; wt = 1.f;      // done_mask, DQN.py:74-77
                }
            }
            float o[8];
            forward_epilogues(bias, o);                             // target net
            float nq = 0.f;
            if (hh == 0) {
                nq = o[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) nq = fmaxf(nq, o[j]);   // q_target(s').max(1), DQN.py:147
            }
            forward_epilogues(bias + 208, o);                       // eval net
            float dbh = 0.f;
            if (hh == 0) {
                float qa = o[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) qa = a_r == j ? o[j] : qa;
                const float y = rew + P.lb.gamma * nq * dmask;      // DQN.py:148, PERDQN.py:163
                const float d = qa - y, ad = fabsf(d);
                float l, g;
                if (MODE == 0) {                                    // smooth-L1, beta = 1, mean over the 32 rows of the event
                    l = (ad < 1.f ? 0.5f * d * d : ad - 0.5f) * wt;
                    g = fminf(fmaxf(d, -1.f), 1.f) * (1.0f / EB) * wt;
                } else {                                            // mean(is_w) * mse (PERDQN.py:182), errors for Memory.update
                    const float ew = e < total ? P.ev_weight[e] : 0.f;
                    l = d * d * wt * ew;
                    g = ew * 2.f * d * (1.0f / EB) * wt;
                    if (e < total) P.lb.new_prio[gr] = ad;
                }
                // per-event loss: a DQN event is the 32 rows of this warp; a PERDQN event the rows of two warps (shared scratch)
                float ls = l;
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o2);
                if (MODE == 0) {
                    if (lane == 0 && e < total) P.lb.loss[e] = ls * (1.0f / EB);
                } else {
                    float* red = reinterpret_cast<float*>(smem + DO_BIAS) + 2 * 208 - 8;          // (last 8 floats of the bias block: unused pad)
                    if (lane == 0) red[q] = ls;
                    head_bar();
                    if (lane == 0 && (q & 1) == 0 && e < total) P.lb.loss[e] = (red[q] + red[q + 1]) * (1.0f / EB);
                    head_bar();
                }
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    w[j] = pk_sat((2 * j == a_r ? g : 0.f) * H_SCALE, (2 * j + 1 == a_r ? g : 0.f) * H_SCALE);
                *reinterpret_cast<uint4*>(sDout + himg(row, 0, 16)) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(sDout + himg(row, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float s = j == a_r ? g : 0.f;
#pragma unroll
                    for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
                    dbh = lane == j ? s : dbh;
                }
            }
            acc_bh += dbh;
            go_signal();                                            // -> dH2, dWh
            wait_done();
            {   // dH2 epilogue: columns [32 hh, +32) of the accumulator in the L2 columns
                const int c0 = hh * 32;
                float v0[32];
                tmem_ld32(T0 + t_lane + TC_L2 + c0, v0);
                tmem_wait_ld();
                mask_store32(v0, sH2, c0, N2);
            }
            go_signal();                                            // -> dW2, dH1, db2
            wait_done();
            {   // dH1 epilogue: columns [64 hh, +64) of the accumulator in the L1 columns
                const int c0 = hh * 64;
                float v0[32], v1[32];
                tmem_ld32(T0 + t_lane + TC_L1 + c0, v0);
                tmem_ld32(T0 + t_lane + TC_L1 + c0 + 32, v1);
                tmem_wait_ld();
                mask_store32(v0, sH1, c0, N1);
                mask_store32(v1, sH1, c0 + 32, N1);
            }
            go_signal();                                            // -> dW1
        }
        if (n_my > 0) {
            // ---- the resident gradients leave TMEM once: plain stores into this CTA's slab ----
            wait_done();
            if (hh == 0 && lane < 8) red_add(G + L::OFF_BH + lane, acc_bh);
            {   // dW1: lane = k1, columns [80 hh, +80) of the 160 inputs, slab layout [x / 4][k1][x % 4]; column 159 = db1
                float* gw = G + L::OFF_W1T + ((hh * 20) * 128 + row) * 4;
#pragma unroll 1
                for (int cb = 0; cb < 5; ++cb) {
                    float v[16];
                    tmem_ld16(T0 + t_lane + TC_DW1 + hh * 80 + cb * 16, v);
                    tmem_wait_ld();
                    if (hh == 1 && cb == 4) { G[L::OFF_B1 + row] = v[15] * (1.0f / H_SCALE); v[15] = 0.f; }
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4)
                        *reinterpret_cast<float4*>(gw + (cb * 4 + j4) * 512) = make_float4(v[j4 * 4] * (1.0f / H_SCALE), v[j4 * 4 + 1] * (1.0f / H_SCALE),
                                                                                            v[j4 * 4 + 2] * (1.0f / H_SCALE), v[j4 * 4 + 3] * (1.0f / H_SCALE));
                }
            }
            {   // dW2: lane = k1, columns [32 hh, +32) of n2 -> W2T slab [k1][n2]
                float v[32];
                tmem_ld32(T0 + t_lane + TC_DW2 + hh * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    *reinterpret_cast<float4*>(G + L::OFF_W2T + row * N2 + hh * 32 + j4 * 4) =
                        make_float4(v[j4 * 4] * (1.0f / H_SCALE), v[j4 * 4 + 1] * (1.0f / H_SCALE), v[j4 * 4 + 2] * (1.0f / H_SCALE), v[j4 * 4 + 3] * (1.0f / H_SCALE));
            }
            if (hh == 0) {   // dWh and db2 (M = 64 accumulators: row n2 = 16 q + i lives in TMEM lane 32 q + i, i < 16)
                float v[16], b2v[16];
                tmem_ld16(T0 + t_lane + TC_DWH, v);
                tmem_ld16(T0 + t_lane + TC_DB2, b2v);
                tmem_wait_ld();
                if (lane < 16) {
                    const int n2 = q * 16 + lane;
#pragma unroll
                    for (int j = 0; j < 8; ++j) G[L::OFF_WH + n2 * 8 + j] = v[j] * (1.0f / H_SCALE);
                    G[L::OFF_B2 + n2] = b2v[0] * (1.0f / H_SCALE);
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(*tmem_slot, 512);
}

// =====================================================================================================
// k_act_dqn_p -- brain.get_action of the DQN-layout brains (DQN.py:126-139, PERDQN.py:101-111) in the same form: 128-row tiles of
// the brain's ALL row list, the net's operand images resident in shared memory, L1 / L2 / head accumulators in disjoint TMEM
// columns (the next tile's L1 runs under this tile's head epilogue), rows gathered float32 -> fp16 by eight warps.  Epilogue =
// k_brain_act's: first-max argmax, exploration draws keyed (t_act, slot) with the family's comparison (coin < eps / u <= eps).
// =====================================================================================================
constexpr int QO_X = 0, QO_H1 = QO_X + XIMG, QO_H2 = QO_H1 + PB * N1 * 2, QO_W = QO_H2 + PB * N2 * 2, QO_BIAS = QO_W + WNET;
constexpr int QO_BARS = QO_BIAS + 4 * 208;
constexpr size_t QACT_SMEM = QO_BARS + 8 * 8 + 16 + 1024;
static_assert(QO_W % 128 == 0 && QO_BARS % 8 == 0, "layout");

struct DqnActParams {
    rl_world_cfg cfg;
    rl_agent_rec* rec;
    const float* obs;          // obs_state
    const int32_t* rows;       // row list of this brain, kind ALL
    const int32_t* total;      // device scalar
    const float* params;
    const double* epsilon;
    uint64_t t_act;
    float* q_out;              // [row_cap][8] or null
    int32_t rule;              // RL_ACT_DQN / RL_ACT_PERDQN
};

__global__ void __launch_bounds__(NTH, 1) k_act_dqn_p(const DqnActParams P) {
    using L = Layout<RL_MODEL_DQN>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __half* sH1 = reinterpret_cast<__half*>(smem + QO_H1);
    __half* sH2 = reinterpret_cast<__half*>(smem + QO_H2);
    float* bias = reinterpret_cast<float*>(smem + QO_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + QO_BARS);
    uint64_t* done = bars; uint64_t* doneL1 = bars + 1; uint64_t* go = bars + 2; uint64_t* xfull = bars + 3; uint64_t* xfree = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = *P.total;
    const int n_tiles = (total + PB - 1) / PB;
    const int n_my = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (threadIdx.x == 0) {
        mbar_init(done, 1); mbar_init(doneL1, 1); mbar_init(go, NEPI); mbar_init(xfull, 32 * NGA); mbar_init(xfree, 1);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    {
        const float* p = P.params;
        __half* w1 = reinterpret_cast<__half*>(smem + QO_W);
        __half* w2 = reinterpret_cast<__half*>(smem + QO_W + WO_W2);
        __half* wh = reinterpret_cast<__half*>(smem + QO_W + WO_WH);
        for (int i = threadIdx.x; i < 160 * N1; i += NTH) { const int k = i / N1, n = i - k * N1; w1[himg(n, k, 160)] = __float2half_rn(p[L::OFF_W1T + i]); }
        for (int i = threadIdx.x; i < N1 * N2; i += NTH) { const int k = i / N2, n = i - k * N2; w2[himg(n, k, N1)] = __float2half_rn(p[L::OFF_W2T + i]); }
        for (int i = threadIdx.x; i < N2 * 16; i += NTH) {
            const int k = i >> 4, j = i & 15;
            wh[himg(j, k, N2)] = __float2half_rn(j < 8 ? p[L::OFF_WH + k * 8 + j] : 0.f);
        }
        for (int i = threadIdx.x; i < 208; i += NTH)
            bias[i] = i < 128 ? p[L::OFF_B1 + i] : i < 192 ? p[L::OFF_B2 + (i - 128)] : i < 200 ? p[L::OFF_BH + (i - 192)] : 0.f;
    }
    fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t aX = smem_u32(smem + QO_X), aH1 = smem_u32(sH1), aH2 = smem_u32(sH2), aW = smem_u32(smem + QO_W);
    constexpr int TA_L2 = 0, TA_L1 = 256, TA_HD = 384;

    if (warp == 9) {
        const bool me = elect_one();
        const uint32_t T0 = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        const int n_u = __shfl_sync(0xffffffffu, n_my, 0);
        uint32_t go_no = 0;
        auto wait_go = [&]() { mbar_wait(go, go_no & 1); ++go_no; fence_after(); };
        auto commit = [&](uint64_t* bar) { if (me) mma_commit(bar); };
        for (int t = 0; t < n_u; ++t) {
            mbar_wait(xfull, t & 1); fence_after();
            {
                const uint32_t id = idesc_h(128, 128, 0, 0);
                uint64_t b = dk(aW, 160);
#pragma unroll 1
                for (int ks = 0; ks < 10; ks += 2) {
                    if (me) { mma_h(T0 + TA_L1, dxk(aX, ks), b, id, ks != 0); mma_h(T0 + TA_L1, dxk(aX, ks + 1), b + 16u, id, 1u); }
                    b += 32u;
                }
                commit(doneL1);
                commit(xfree);
            }
            wait_go();
            {
                const uint32_t id = idesc_h(128, N2, 0, 0);
                uint64_t a = dk(aH1, N1), b = dk(aW + WO_W2, N1);
#pragma unroll 1
                for (int ks = 0; ks < 8; ks += 4) {
                    if (me) {
                        mma_h(T0 + TA_L2, a, b, id, ks != 0); mma_h(T0 + TA_L2, a + 16u, b + 16u, id, 1u);
                        mma_h(T0 + TA_L2, a + 32u, b + 32u, id, 1u); mma_h(T0 + TA_L2, a + 48u, b + 48u, id, 1u);
                    }
                    a += 64u; b += 64u;
                }
                commit(done);
            }
            wait_go();
            {
                const uint32_t id = idesc_h(128, 16, 0, 0);
                const uint64_t a = dk(aH2, N2), b = dk(aW + WO_WH, N2);
                if (me) {
                    mma_h(T0 + TA_HD, a, b, id, 0u); mma_h(T0 + TA_HD, a + 16u, b + 16u, id, 1u);
                    mma_h(T0 + TA_HD, a + 32u, b + 32u, id, 1u); mma_h(T0 + TA_HD, a + 48u, b + 48u, id, 1u);
                }
                commit(done);
            }
        }
    } else if (warp >= 10) {
        const int gt = threadIdx.x - 320;
        for (int t = 0; t < n_my; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            float4 xa[10], xb[10];
#pragma unroll
            for (int u = 0; u < 10; ++u) {
                const int v = gt + u * (32 * NGA);
                const int r = v / 20, oct = v - r * 20;
                const int i = tile * PB + r;
                const int rid = i < total ? __ldg(P.rows + i) : 0;
                const float4* g = reinterpret_cast<const float4*>(P.obs + (size_t)rid * RL_K1) + oct * 2;
                xa[u] = ld_stream_f4(g); xb[u] = ld_stream_f4(g + 1);
            }
            if (t > 0) mbar_wait(xfree, (t - 1) & 1);
#pragma unroll
            for (int u = 0; u < 10; ++u) {
                const int v = gt + u * (32 * NGA);
                const int r = v / 20, oct = v - r * 20;
                *reinterpret_cast<uint4*>(smem + QO_X + ximg(r, oct)) =
                    make_uint4(pk(xa[u].x, xa[u].y), pk(xa[u].z, xa[u].w), pk(xb[u].x, xb[u].y), pk(xb[u].z, xb[u].w));
            }
            fence_proxy_async();
            mbar_arrive(xfull);
        }
    } else if (warp < 8) {
        const uint32_t T0 = *tmem_slot;
        uint32_t done_no = 0, l1_no = 0;
        const int q = warp & 3, hh = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const int S = P.cfg.slot_cap;
        auto go_signal = [&]() { fence_proxy_async(); fence_before(); mbar_arrive(go); };
        auto wait_done = [&]() { mbar_wait(done, done_no & 1); ++done_no; fence_after(); };
        auto relu_store32 = [&](float (&v)[32], const float* b, __half* img, int c0, int K) {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                const float4 b0 = *reinterpret_cast<const float4*>(b + j8 * 8), b1 = *reinterpret_cast<const float4*>(b + j8 * 8 + 4);
                *reinterpret_cast<uint4*>(img + himg(row, c0 + j8 * 8, K)) =
                    make_uint4(pk_relu(v[j8 * 8] + b0.x, v[j8 * 8 + 1] + b0.y), pk_relu(v[j8 * 8 + 2] + b0.z, v[j8 * 8 + 3] + b0.w),
                               pk_relu(v[j8 * 8 + 4] + b1.x, v[j8 * 8 + 5] + b1.y), pk_relu(v[j8 * 8 + 6] + b1.z, v[j8 * 8 + 7] + b1.w));
            }
        };
        const double epsilon = *P.epsilon;
        for (int t = 0; t < n_my; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const int i = tile * PB + row;
            const int rid = (hh == 0 && i < total) ? __ldg(P.rows + i) : 0;
            mbar_wait(doneL1, l1_no & 1); ++l1_no; fence_after();
            {
                const int c0 = hh * 64;
                float v0[32], v1[32];
                tmem_ld32(T0 + t_lane + TA_L1 + c0, v0);
                tmem_ld32(T0 + t_lane + TA_L1 + c0 + 32, v1);
                tmem_wait_ld();
                relu_store32(v0, bias + c0, sH1, c0, N1);
                relu_store32(v1, bias + c0 + 32, sH1, c0 + 32, N1);
            }
            go_signal();
            wait_done();
            {
                const int c0 = hh * 32;
                float v0[32];
                tmem_ld32(T0 + t_lane + TA_L2 + c0, v0);
                tmem_wait_ld();
                relu_store32(v0, bias + 128 + c0, sH2, c0, N2);
            }
            go_signal();
            wait_done();
            if (hh == 0) {
                float v[16];
                tmem_ld16(T0 + t_lane + TA_HD, v);
                tmem_wait_ld();
                if (i < total) {
                    float qv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) qv[j] = v[j] + bias[192 + j];
                    int best = 0;
#pragma unroll
                    for (int j = 1; j < 8; ++j) if (qv[j] > qv[best]) best = j;                   // first maximum
                    const int w = rid / S, slot = rid - w * S;
                    const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));
                    int a = best;
                    const double u = rl_uniform(rl_draw(key, P.t_act, RL_SITE_ACT_EXPLORE, (uint32_t)slot));
                    const bool explore = P.rule == RL_ACT_DQN ? (u < epsilon) : (u <= epsilon);    // DQN.py:135-139 / PERDQN.py:101-111
                    if (explore) a = (int)rl_below(rl_draw(key, P.t_act, RL_SITE_ACT_RANDOM, (uint32_t)slot), 8);
                    reinterpret_cast<int8_t*>(P.rec + rid)[13] = (int8_t)a;
                    if (P.q_out) {
                        float4* qo = reinterpret_cast<float4*>(P.q_out + (size_t)i * 8);
                        qo[0] = make_float4(qv[0], qv[1], qv[2], qv[3]);
                        qo[1] = make_float4(qv[4], qv[5], qv[6], qv[7]);
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(*tmem_slot, 512);
}

int launch(int mode, const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
           const int32_t* sample_idx, const float* ev_weight, const rl_learn_bufs* learn, void* stream) {
    DqnParams P;
    memset(&P, 0, sizeof(P));
    P.cfg = *cfg;
    P.ev_rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_EVENT) * rows->row_cap;
    P.ev_total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    P.rp = *replay; P.sample_idx = sample_idx; P.ev_weight = ev_weight; P.lb = *learn;
    static PerDeviceOnce attr;
    if (attr.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_learn_dqn_p<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DQP_SMEM));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_learn_dqn_p<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DQP_SMEM));
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) k_learn_dqn_p<0><<<rl_learn_grid(), NTH, DQP_SMEM, st>>>(P);
    else k_learn_dqn_p<1><<<rl_learn_grid(), NTH, DQP_SMEM, st>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    int rc = rl_learn_reduce(learn, P.ev_total, 2, stream);
    if (rc) return rc;
    return rl_count_valid_launch(sample_idx, mode == 0 ? 32 : 64, P.ev_total, learn->grad + Layout<RL_MODEL_DQN>::N_TRAIN, stream);
}

}  // namespace

extern "C" int rl_brain_learn_dqn_p(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                                    const int32_t* sample_idx, const rl_learn_bufs* learn, void* stream) {
    if (replay && replay->obs_fp16) return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_learn_dqn_p: float16 replay rows are not supported");
    RL_ARG_CHECK(cfg && rows && replay && sample_idx && learn);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1);
    RL_ARG_CHECK(learn->kind == RL_MODEL_DQN && learn->batch == 32);
    RL_ARG_CHECK(learn->params && learn->target && learn->grad_scratch && learn->grad && learn->loss);
    RL_ARG_CHECK((int64_t)cfg->n_worlds * replay->capacity < (1ll << 31));
    return launch(0, cfg, rows, gene, replay, sample_idx, nullptr, learn, stream);
}

extern "C" int rl_brain_learn_perdqn_p(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                                       const int32_t* sample_idx, const float* ev_weight, const rl_learn_bufs* learn, void* stream) {
    if (replay && replay->obs_fp16) return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_learn_perdqn_p: float16 replay rows are not supported");
    RL_ARG_CHECK(cfg && rows && replay && sample_idx && ev_weight && learn);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1);
    RL_ARG_CHECK(learn->kind == RL_MODEL_DQN && learn->batch == 64);
    RL_ARG_CHECK(learn->params && learn->target && learn->grad_scratch && learn->grad && learn->loss && learn->new_prio);
    RL_ARG_CHECK((int64_t)cfg->n_worlds * replay->capacity < (1ll << 31));
    return launch(1, cfg, rows, gene, replay, sample_idx, ev_weight, learn, stream);
}

extern "C" int rl_brain_act_dqn_p(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                                  const rl_brain_act* brain, uint64_t t_act, float* q_out, void* stream) {
    RL_ARG_CHECK(cfg && bufs && rows && brain);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1);
    RL_ARG_CHECK(bufs->rec && bufs->obs_state && brain->params && brain->epsilon);
    if (brain->kind != RL_MODEL_DQN || (brain->rule != RL_ACT_DQN && brain->rule != RL_ACT_PERDQN))
        return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_act_dqn_p: DQN-layout networks (DQN, PERDQN) only");
    DqnActParams P;
    memset(&P, 0, sizeof(P));
    P.cfg = *cfg; P.rec = bufs->rec; P.obs = bufs->obs_state;
    P.rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_ALL) * rows->row_cap;
    P.total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_ALL;
    P.params = brain->params; P.epsilon = brain->epsilon; P.t_act = t_act; P.rule = brain->rule;
    P.q_out = q_out ? q_out + (size_t)gene * rows->row_cap * 8 : nullptr;
    static PerDeviceOnce attr;
    if (attr.need()) RL_CUDA_CHECK(cudaFuncSetAttribute(k_act_dqn_p, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QACT_SMEM));
    k_act_dqn_p<<<rl_learn_grid(), NTH, QACT_SMEM, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

// tc_act_kernels.cu -- k_act_dueling_p: brain.get_action of the dueling brains (PERD3QN.py:197-210) on the 5th-gen tensor cores in
// the BATCH-MAJOR form of the paired event kernel (tc_pair_kernels.cu, tc_bm.cuh): a tile is 128 rows of the brain's ALL row list,
//   L1  D[128 b][128]  = X[128][160] W1^T      X: SWIZZLE_128B image, gathered from obs_state (float32 -> fp16) by eight gather warps
//   L2  D[128 b][256]  = H1[128][128] W2^T     H1 / H2: interleaved no-swizzle images written by the epilogue (relu(D + b), F2FP.RELU)
//   hd  D[128 b][16]   = H2[128][256] Wh^T     per-row dueling combine at B = 1, first-max argmax, exploration draws keyed (t_act, slot)
// One warp-uniform MMA issuer (elect.sync), weight chunks streamed through an 8 x 8 KB ring with one barrier per release group, the
// three accumulators in disjoint TMEM columns (L2 0..255, L1 256..383, head 384..399) so that the next tile's L1 runs under this
// tile's head epilogue; the next tile's rows are in the gather warps' registers while this tile computes (their loads are issued
// before the X image is free).  Replaces k_act_dueling_h (transposed-output form, N = 64 MMAs, scalar F2F conversions and 2-byte
// scatter stores: 0.12 ms per 200 k rows) -- same contract, outputs and tolerance (tests/test_tc_gpu.py, tests/test_scale_gpu.py).
// TMA = true (rl_world_bufs.obs_state_h set: the World kernels keep float16 copies of the rows): no register gather -- four warps issue
// 96 cp.async.bulk.tensor tile::gather4 per tile by row id from a [n_worlds * slot_cap][160] tensor map straight into the swizzled image
// (no LSU traffic beside the MMAs: 0.072 -> 0.045 ms per launch; bit-identical results, same float16 operands).
#include <stdlib.h>
#include <string.h>
#include "tc_bm.cuh"
#include "tma_util.cuh"
#include "models.cuh"

namespace {

using namespace tc;
using namespace bm;
using mlp::mbar_init; using mlp::mbar_wait; using mlp::fence_mbar_init; using mlp::fence_proxy_async;

constexpr int NEPI = 256;                // epilogue threads (warps 0-7)
constexpr int NGA = 8;                   // gather warps 10..17
constexpr int NTH = NEPI + 64 + 32 * NGA;    // + weight producer (warp 8) + MMA issuer (warp 9)
constexpr int NSP = 8;                   // weight-chunk ring slots of 8 KB
constexpr int WI_W1 = 0, WI_W2K = 5, WI_WH = 13;      // weight image chunk ids (tc_kernels.cu::k_build_wimg_dueling_h)
constexpr int SCHED_A = 14;              // chunks per tile: W1[5] W2K[8] WH
constexpr int NSTG = 4;                  // release groups per tile: L1 (5) L2a (4) L2b (4) head (1)

constexpr int AO_X = 0;                                   // X, SWIZZLE_128B
constexpr int AO_H2 = AO_X + XIMG;                        // H2 [128][256]
constexpr int AO_H1 = AO_H2 + PB * 256 * 2;               // H1 [128][128]
constexpr int AO_STG = AO_H1 + PB * 128 * 2;              // weight ring
constexpr int AO_BIAS = AO_STG + NSP * HCH * 2;           // b1[128] b2[256] bh[16] floats
constexpr int AO_BARS = AO_BIAS + 4 * 400;
constexpr int NBAR = 2 * NSTG + 5;       // gfull[NSTG] sfree[NSTG] done doneL1 go xfull xfree
constexpr size_t ACT_SMEM = AO_BARS + 8 * NBAR + 16 + 1024;
static_assert(ACT_SMEM <= 227 * 1024 && AO_BARS % 8 == 0 && AO_H2 % 1024 == 0, "shared memory budget");

__device__ __forceinline__ uint32_t group_chunks(uint32_t g) { return g == 0 ? 5u : g == 3 ? 1u : 4u; }

struct ActParams {
    CUtensorMap map_obs;       // TMA = true: obs_state_h as a [n_worlds * slot_cap][160] float16 tensor
    rl_world_cfg cfg;
    rl_agent_rec* rec;
    const float* obs;          // obs_state
    const int32_t* rows;       // row list of this brain, kind ALL
    const int32_t* total;      // device scalar
    const float* params;       // biases are read from the kernel-layout buffer
    const __half* wimg;
    const double* epsilon;
    uint64_t t_act;
    float* q_out;              // [row_cap][8] or null
    long long* trace;          // RL_TC_TRACE: clock64 stamps of CTA 0, tile 3
};

template <bool TMA>
__global__ void __launch_bounds__(NTH, 1) k_act_dueling_p(const __grid_constant__ ActParams P) {
    using L = Layout<RL_MODEL_DUELING>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __half* sH1 = reinterpret_cast<__half*>(smem + AO_H1);
    __half* sH2 = reinterpret_cast<__half*>(smem + AO_H2);
    __half* sStg = reinterpret_cast<__half*>(smem + AO_STG);
    float* bias = reinterpret_cast<float*>(smem + AO_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AO_BARS);
    uint64_t* gfull = bars; uint64_t* sfree = bars + NSTG; uint64_t* done = sfree + NSTG; uint64_t* doneL1 = done + 1; uint64_t* go = done + 2;
    uint64_t* xfull = done + 3; uint64_t* xfree = done + 4;     // X image written by the gather warps / read out by the L1 MMAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = *P.total;
    const int n_tiles = (total + PB - 1) / PB;
    const int n_my = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTG; ++i) { mbar_init(&gfull[i], 1); mbar_init(&sfree[i], 1); }
        mbar_init(done, 1); mbar_init(doneL1, 1); mbar_init(go, NEPI); mbar_init(xfull, TMA ? 4 : 32 * NGA); mbar_init(xfree, 1);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < NEPI)
        for (int i = threadIdx.x; i < 400; i += NEPI) {
            const int o = i < 128 ? L::OFF_B1 + i : i < 384 ? L::OFF_B2 + (i - 128) : L::OFF_BH + (i - 384);
            bias[i] = i < 384 + 9 ? P.params[o] : 0.f;
        }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t aX = smem_u32(smem + AO_X), aH1 = smem_u32(sH1), aH2 = smem_u32(sH2), aStg = smem_u32(sStg);

    if (warp == 8) {
        // =================================== weight-stream producer ===================================
        if (lane == 0) {
            const uint32_t n_chunks = (uint32_t)n_my * SCHED_A;
            uint32_t freed = 0, stage = 0, grp = 0, left = 0;
            for (uint32_t produced = 0; produced < n_chunks; ++produced) {
                while (produced - freed >= (uint32_t)NSP) {
                    const uint32_t k = stage % NSTG;
                    mbar_wait(&sfree[k], (stage / NSTG) & 1);
                    freed += group_chunks(k);
                    ++stage;
                }
                if (left == 0) {
                    left = group_chunks(grp);
                    fence_proxy_async();
                    mbar_expect_tx(&gfull[grp], left * (uint32_t)(HCH * 2));
                }
                const uint32_t i = produced % SCHED_A;
                const int ch = i < 5 ? WI_W1 + (int)i : i < 13 ? WI_W2K + (int)(i - 5) : WI_WH;
                bulk_copy(sStg + (produced % NSP) * HCH, P.wimg + (size_t)ch * HCH, HCH * 2, &gfull[grp]);
                if (--left == 0) grp = grp + 1 == NSTG ? 0 : grp + 1;
            }
        }
    } else if (warp == 9) {
        // =================================== MMA issuer (one warp, warp-uniform) ===================================
        const bool me = elect_one();
        const uint32_t T0 = __shfl_sync(0xffffffffu, *tmem_slot, 0);
        const int n_u = __shfl_sync(0xffffffffu, n_my, 0);
        uint32_t consumed = 0, go_no = 0, stage = 0, wstage = 0;
        int itr_n = 0, itr_t = -1;
        auto istamp = [&]() { if (P.trace && blockIdx.x == 0 && me && itr_t == 3 && itr_n < 30) P.trace[32 + itr_n++] = clock64(); };
        auto wait_go = [&]() { mbar_wait(go, go_no & 1); ++go_no; fence_after(); istamp(); };
        auto chunks_wait = [&](int n) -> uint32_t {
            const uint32_t first = consumed;
            mbar_wait(&gfull[wstage % NSTG], (wstage / NSTG) & 1);
            ++wstage;
            consumed += n;
            return first;
        };
        auto chunk_addr = [&](uint32_t k) -> uint32_t { return aStg + (k % NSP) * (HCH * 2); };
        auto commit = [&](uint64_t* bar) { if (me) mma_commit(bar); };
        auto stage_free = [&]() { commit(&sfree[stage % NSTG]); ++stage; };
        for (int t = 0; t < n_u; ++t) {
            itr_t = t; istamp();
            uint32_t k0 = chunks_wait(5);
            istamp();
            mbar_wait(xfull, t & 1);
            fence_after();
            istamp();
            {   // L1: 5 chunks [128 n][32 k], 2 k-steps each -> columns 256..383
                const uint32_t id = idesc_h(128, 128, 0, 0);
#pragma unroll 1
                for (int c = 0; c < 5; ++c) {
                    const uint64_t b = dk(chunk_addr(k0 + c), 32);
                    if (me) {
                        mma_h(T0 + 256, dxk(aX, 2 * c), b, id, c != 0);
                        mma_h(T0 + 256, dxk(aX, 2 * c + 1), b + 16u, id, 1u);
                    }
                }
                istamp();
                stage_free();
                commit(doneL1);
                commit(xfree);                                 // the X image may be overwritten with the next tile's rows
            }
            k0 = chunks_wait(4);
            wait_go();
            {   // L2: 8 chunks [256 n][16 k] -> columns 0..255
                const uint32_t id = idesc_h(128, 256, 0, 0);
                uint64_t a = dk(aH1, 128);
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    if (hf) k0 = chunks_wait(4);
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        if (me) mma_h(T0, a, dk(chunk_addr(k0 + c), 16), id, (hf | c) != 0);
                        a += 16u;
                    }
                    istamp();
                    stage_free();
                }
                commit(done);
            }
            k0 = chunks_wait(1);
            wait_go();
            {   // head: one chunk [16 n][256 k], 16 k-steps -> columns 384..399
                const uint32_t id = idesc_h(128, 16, 0, 0);
                uint64_t a = dk(aH2, 256), b = dk(chunk_addr(k0), 256);
#pragma unroll 1
                for (int ks = 0; ks < 16; ks += 4) {
                    if (me) {
                        mma_h(T0 + 384, a, b, id, ks != 0); mma_h(T0 + 384, a + 16u, b + 16u, id, 1u);
                        mma_h(T0 + 384, a + 32u, b + 32u, id, 1u); mma_h(T0 + 384, a + 48u, b + 48u, id, 1u);
                    }
                    a += 64u; b += 64u;
                }
                istamp();
                stage_free();
                commit(done);
            }
        }
    } else if (warp >= 10) {
        // =================================== row gatherers (eight warps) ===================================
        // 128 rows x 20 units of 8 float32 columns -> packed fp16, 10 units per thread; a quarter-warp takes 8 consecutive units of
        // a row (256 contiguous bytes, conflict-free swizzled stores).  The loads of tile t + 1 are issued before the X image is
        // free (xfree: the L1 MMAs of tile t have completed), so their latency sits behind tile t.  Rows past `total` read row 0.
        const int gt = threadIdx.x - 320;
        if (TMA) {
            // float16 copies of the rows exist (rl_world_bufs.obs_state_h): the TMA gathers them by row id straight into the
            // SWIZZLE_128B image -- 32 row groups x 3 K blocks = 96 tile::gather4 over four warps, no LSU traffic beside the MMAs
            const int gw = warp - 10;
            if (gw < 4) {
                const int g = (lane / 3) * 4 + gw, kb = lane - (lane / 3) * 3;
                const bool act = lane < 24;
                if (lane == 0) tma::prefetch_map(&P.map_obs);
                for (int t = 0; t < n_my; ++t) {
                    const int tile = (int)blockIdx.x + t * (int)gridDim.x;
                    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
                    if (act) {
                        const int i = tile * PB + 4 * g;
                        r0 = i < total ? __ldg(P.rows + i) : 0; r1 = i + 1 < total ? __ldg(P.rows + i + 1) : 0;
                        r2 = i + 2 < total ? __ldg(P.rows + i + 2) : 0; r3 = i + 3 < total ? __ldg(P.rows + i + 3) : 0;
                    }
                    if (t > 0) mbar_wait(xfree, (t - 1) & 1);
                    if (lane == 0) mbar_expect_tx(xfull, XIMG / 4);
                    __syncwarp();
                    if (act) tma::gather4(smem_u32(smem + AO_X) + kb * XBLK + g * 512, &P.map_obs, smem_u32(xfull), kb * 64, r0, r1, r2, r3);
                }
            }
        } else
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
            if (P.trace && blockIdx.x == 0 && gt == 0 && t == 4) P.trace[64] = clock64();
            if (t > 0) mbar_wait(xfree, (t - 1) & 1);
            if (P.trace && blockIdx.x == 0 && gt == 0 && t == 4) P.trace[65] = clock64();
#pragma unroll
            for (int u = 0; u < 10; ++u) {
                const int v = gt + u * (32 * NGA);
                const int r = v / 20, oct = v - r * 20;
                *reinterpret_cast<uint4*>(smem + AO_X + ximg(r, oct)) =
                    make_uint4(pk(xa[u].x, xa[u].y), pk(xa[u].z, xa[u].w), pk(xb[u].x, xb[u].y), pk(xb[u].z, xb[u].w));
            }
            fence_proxy_async();
            mbar_arrive(xfull);
            if (P.trace && blockIdx.x == 0 && gt == 0 && t == 4) P.trace[66] = clock64();
        }
    } else {
        // =================================== epilogue warps ===================================
        const uint32_t T0 = *tmem_slot;
        uint32_t done_no = 0, l1_no = 0;
        const int q = warp & 3, hh = warp >> 2;
        const int row = q * 32 + lane;                    // batch row of the tile == TMEM lane
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const int S = P.cfg.slot_cap;
        int tr_n = 0, tr_t = -1;
        auto stamp = [&]() { if (P.trace && blockIdx.x == 0 && threadIdx.x == 0 && tr_t == 3 && tr_n < 30) P.trace[tr_n++] = clock64(); };
        auto go_signal = [&]() { fence_proxy_async(); fence_before(); mbar_arrive(go); stamp(); };
        auto wait_done = [&]() { mbar_wait(done, done_no & 1); ++done_no; fence_after(); stamp(); };
        auto wait_l1 = [&]() { mbar_wait(doneL1, l1_no & 1); ++l1_no; fence_after(); stamp(); };
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
            tr_t = t; stamp();
            wait_l1();
            {   // L1 epilogue: this thread's row, columns [64 hh, +64) of the accumulator at TMEM columns 256..383
                const int c0 = hh * 64;
                float v0[32], v1[32];
                tmem_ld32(T0 + t_lane + 256 + c0, v0);
                tmem_ld32(T0 + t_lane + 256 + c0 + 32, v1);
                tmem_wait_ld();
                relu_store32(v0, bias + c0, sH1, c0, 128);
                relu_store32(v1, bias + c0 + 32, sH1, c0 + 32, 128);
            }
            go_signal();                                            // -> L2
            wait_done();
#pragma unroll 1
            for (int cb = 0; cb < 2; ++cb) {                        // L2 epilogue: columns [128 hh, +128)
                const int c0 = hh * 128 + cb * 64;
                float v0[32], v1[32];
                tmem_ld32(T0 + t_lane + c0, v0);
                tmem_ld32(T0 + t_lane + c0 + 32, v1);
                tmem_wait_ld();
                relu_store32(v0, bias + 128 + c0, sH2, c0, 256);
                relu_store32(v1, bias + 128 + c0 + 32, sH2, c0 + 32, 256);
            }
            go_signal();                                            // -> head (the next tile's L1 follows it)
            wait_done();
            if (hh == 0) {
                float v[16];
                tmem_ld16(T0 + t_lane + 384, v);
                tmem_wait_ld();
                if (i < total) {
                    float qv[8], ssum = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { qv[j] = v[j] + bias[384 + j]; ssum += qv[j]; }
                    const float val = v[8] + bias[384 + 8], mean = ssum * 0.125f;                 // B = 1: per-row mean (PERD3QN.py:202)
#pragma unroll
                    for (int j = 0; j < 8; ++j) qv[j] = qv[j] + val - mean;
                    int best = 0;
#pragma unroll
                    for (int j = 1; j < 8; ++j) if (qv[j] > qv[best]) best = j;                   // first maximum
                    const int w = rid / S, slot = rid - w * S;
                    const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));
                    int a = best;                                                                  // PERD3QN.py:204-210
                    const double u = rl_uniform(rl_draw(key, P.t_act, RL_SITE_ACT_EXPLORE, (uint32_t)slot));
                    if (!(u > epsilon)) a = (int)rl_below(rl_draw(key, P.t_act, RL_SITE_ACT_RANDOM, (uint32_t)slot), 8);
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

}  // namespace

extern "C" int rl_brain_act_p(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                              const rl_brain_act* brain, const void* wimg_eval_h, uint64_t t_act, float* q_out, void* stream) {
    RL_ARG_CHECK(cfg && bufs && rows && brain && wimg_eval_h);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1);
    RL_ARG_CHECK(bufs->rec && bufs->obs_state && brain->params && brain->epsilon);
    if (brain->kind != RL_MODEL_DUELING || brain->rule != RL_ACT_DUELING)
        return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_act_p: dueling networks only");
    ActParams P;
    memset(&P, 0, sizeof(P));
    P.cfg = *cfg; P.rec = bufs->rec; P.obs = bufs->obs_state;
    P.rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_ALL) * rows->row_cap;
    P.total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_ALL;
    P.params = brain->params; P.wimg = reinterpret_cast<const __half*>(wimg_eval_h); P.epsilon = brain->epsilon; P.t_act = t_act;
    P.q_out = q_out ? q_out + (size_t)gene * rows->row_cap * 8 : nullptr;
    static long long* trace_dev = nullptr;
    const bool tracing = getenv("RL_TC_TRACE") != nullptr;
    if (tracing) {
        if (!trace_dev) RL_CUDA_CHECK(cudaMalloc(&trace_dev, 128 * sizeof(long long)));
        RL_CUDA_CHECK(cudaMemset(trace_dev, 0, 128 * sizeof(long long)));
        P.trace = trace_dev;
    }
    const bool use_tma = bufs->obs_state_h != nullptr && getenv("RL_ACT_NO_TMA") == nullptr;
    if (use_tma) {
        const long long n_rows = (long long)cfg->n_worlds * cfg->slot_cap;
        RL_ARG_CHECK(n_rows < (1ll << 31));
        if (tma::make_rows_map(&P.map_obs, bufs->obs_state_h, (uint64_t)n_rows, RL_K1, 1) != 0)
            return rl_set_err(RL_ERR_CUDA, "rl_brain_act_p: cuTensorMapEncodeTiled failed");
    }
    static PerDeviceOnce attr;
    if (attr.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_act_dueling_p<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ACT_SMEM));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_act_dueling_p<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ACT_SMEM));
    }
    if (use_tma) k_act_dueling_p<true><<<rl_learn_grid(), NTH, ACT_SMEM, (cudaStream_t)stream>>>(P);
    else k_act_dueling_p<false><<<rl_learn_grid(), NTH, ACT_SMEM, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    if (tracing) {
        long long h[128];
        RL_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
        RL_CUDA_CHECK(cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[act trace] epilogue stamps (tile start, L1 done, ->L2, L2 done, ->head, head done):");
        for (int i = 1; i < 30 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
        fprintf(stderr, "\n[act trace] issuer stamps (tile start, W1 chunks, xfull, L1 issued, go, L2a issued, L2b issued, go, head issued):");
        for (int i = 0; i < 30 && h[32 + i]; ++i) fprintf(stderr, " %lld", h[32 + i] - h[0]);
        fprintf(stderr, "\n[act trace] gather warps, tile 4 (loads issued, xfree seen, stored):");
        for (int i = 0; i < 3; ++i) fprintf(stderr, " %lld", h[64 + i] - h[0]);
        fprintf(stderr, "\n");
    }
    return RL_OK;
}

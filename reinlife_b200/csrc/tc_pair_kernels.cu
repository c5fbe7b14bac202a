// tc_pair_kernels.cu -- k_learn_dueling_p: the train() events of the dueling brains (PERD3QN.py:94-115, D3QN.py:97-116)
// on the 5th-gen tensor cores, TWO events (2 x 64 rows) per CTA iteration in BATCH-MAJOR form.
//
// Why a second design next to tc_kernels.cu::k_learn_dueling_h.  An event processed alone (batch 64) is issue-bound: the
// transposed-output form of k_learn_dueling_h reaches M = 128 but every MMA is N = 64, every operand needs a batch-major AND a
// feature-major image (scalar scatter stores), and every stage is handed over once per 64 rows.  Here
//   * a pair of events is one M = 128 tile: D[batch 128][features] = Act[128][K] * W[features][K]^T, N = 128 / 256 per MMA;
//   * fp16 operands may be MN-major (scripts/tc_probe_h.py pins it): the weight-gradient GEMMs dW2 = H1^T dH2,
//     dW1 = dH1^T X, dWh = H2^T dOut read the SAME batch-major images the forward wrote, so no transposed image exists:
//     X, H1, H2, dOut only, dH2 / dH1 written in place;
//   * an epilogue thread owns one batch ROW: activations leave as 16-byte vector stores, the dueling combine / TD
//     error / dOut of a row are thread-local, bias gradients are warp reduce-scatters (db2, dbh) or ride a GEMM for free
//     (db1 = column 159 of dW1, whose X column is 1 -- the matching W1 rows are structural zeros);
//   * dW2 stays resident in TMEM across all pairs of the CTA, dW1 is flushed once per pair with 16-byte vector reds.
// Round-2 rework (what scripts/tc_issue_probe.py, scripts/tc_probe_sw128.py and the RL_TC_TRACE stamps showed):
//   * ONE issuer warp in warp-uniform control flow + elect.sync: descriptors stay in uniform registers and UTCHMMA issues
//     directly (48-69 cycles per N = 64-128 MMA, the math time at N = 256).  Picking the issuing thread with `lane == 0` makes
//     ptxas wrap every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (92 cycles per MMA whatever its N), and two such
//     issuers only shared the tensor pipe;
//   * the sampled replay rows are gathered by the TMA (cp.async.bulk.tensor tile::gather4, four ring rows per instruction)
//     straight into SWIZZLE_128B operand images: 128 rows x 160 halves = three [128][128 B] K blocks.  The same image is the
//     K-major A operand of L1 and the MN-major B operand of dW1 (the XOR pattern is the same in both readings).  The LSU
//     gather it replaces (2560 16-byte loads per image, 8 rows per quarter-warp) took 3-6 k cycles to drain and stalled the
//     MMA stages it was meant to hide behind (tL2 6.5 k -> 2.2 k cycles, dH1 8.3 k -> 3.3 k without it);
//   * the target rows X' of the NEXT pair land in the H2 region (dead once dW2 / dH1 have completed), the eval rows X in the X
//     region (dead once dW1 has completed): both gathers are off the epilogue's critical path and need no registers.
// float32 rings (rl_replay_bufs.obs_fp16 = 0) cannot be gathered by the TMA (it does not convert): they keep a register gather
// that writes the same swizzled images (template parameter TMA = false).
// Same contract, outputs and tolerance as rl_brain_learn_h (tests/test_scale_gpu.py, tests/test_tc_gpu.py).
#include <stdlib.h>
#include <string.h>
#include <cuda_fp16.h>
#include "tc_tile.cuh"
#include "tma_util.cuh"
#include "tc_bm.cuh"
#include "models.cuh"

namespace {

using namespace tc;
using namespace bm;
using mlp::mbar_init; using mlp::mbar_wait; using mlp::fence_mbar_init; using mlp::fence_proxy_async; using mlp::bulk_load;

constexpr int R = 64;                    // rows per event
constexpr int NEPI = 256;                // epilogue threads (warps 0-7)
constexpr int NTH = NEPI + 64;           // + weight producer (warp 8) + MMA issuer (warp 9)
constexpr int NGW = 4;                   // TMA = true: + warps 10, 11 (row gatherers: they complete warpgroup 2) + NGW accumulator / row-gatherer warps 12..
constexpr int NGI = NGW + 2;             // warps that issue the row gathers of an image: 96 tile::gather4 = NGI x 16 lanes
constexpr int NTH_TMA = NEPI + 128 + 32 * NGW;
// register budget of the TMA variant (setmaxnreg, per warpgroup of 4 warps): 512 threads are launched with 128 registers each (the pool
// setmaxnreg redistributes is the CTA's launch allocation); the epilogue warpgroups 0-1 keep them, warpgroup 2 (producer, issuer) shrinks
// to REG_AUX, the accumulator warpgroup 3 grows to REG_ACC
constexpr int REG_EPI = 128, REG_AUX = 40, REG_ACC = 216;
static_assert(256 * REG_EPI + 128 * REG_AUX + 32 * NGW * REG_ACC <= NTH_TMA * 128, "register pool of the launch");
constexpr int NSP = 8;                   // weight-chunk ring slots of 8 KB
constexpr float H_SCALE = 256.0f;        // backward operands are scaled by 2^8 (exact), removed when gradients leave TMEM

// weight image chunk ids (tc_kernels.cu::k_build_wimg_dueling_h)
constexpr int WI_W1 = 0, WI_WH = 13, WI_WHT = 14, WI_W2T = 15, WI_W2N = 23;    // (W2N: L2 weights split by output halves, 4 chunks [128 n][32 k] each)
constexpr int SCHED_P = 37;              // chunks per pair: t W1[5] W2N[8] WH | e W1[5] W2N[8] WH WHT W2T[8]

// ---- shared memory (bytes, from a 1024-byte aligned base) ----
constexpr int PO_X = 0;                                   // X (eval rows), SWIZZLE_128B
constexpr int PO_H2 = PO_X + XIMG;                        // H2 -> dH2 [128][256] no swizzle; between pairs: X' (target rows), SWIZZLE_128B
constexpr int PO_H1 = PO_H2 + PB * 256 * 2;               // H1 -> dH1 [128][128]
constexpr int PO_STG = PO_H1 + PB * 128 * 2;              // weight ring
constexpr int PO_DOUT = PO_STG + NSP * HCH * 2;           // dOut [128][16] halves
constexpr int PO_BIAS = PO_DOUT + PB * 16 * 2;            // b1[128] b2[256] bh[16] x {target, eval} floats
constexpr int PO_RED = PO_BIAS + 4 * 800;                 // reduction scratch floats [64]
constexpr int PO_META = PO_RED + 4 * 64;                  // [2] x { idx[128] act[128] rew[128] dn[128] } int / float
constexpr int PO_RING = PO_META + 2 * 4 * 4 * PB;         // [2][2] ring base (elements) per buffer / event, 64-bit
constexpr int PO_BARS = PO_RING + 2 * 2 * 8;
constexpr int NSTG = 11;                 // ring release groups per pair: tL1 tL2a tL2b thead eL1 eL2a eL2b ehead dH2 dH1a dH1b
constexpr int NBAR = 2 * NSTG + 12;      // gfull[NSTG] sfree[NSTG] + done doneL1 go xfull xpfull h2free xfree metaready w1a w1b done2 go2
constexpr int PO_ONES = (PO_BARS + 8 * NBAR + 16 + 127) & ~127;   // [16 k][16] halves of 1.0: B operand of the db2 GEMM (every k-step reads it)
constexpr size_t PAIR_SMEM = PO_ONES + 512 + 1024;               // + alignment slack
static_assert(PAIR_SMEM <= 227 * 1024 && PO_BARS % 8 == 0 && PO_RING % 8 == 0 && PO_H2 % 1024 == 0, "shared memory budget");

__device__ __forceinline__ uint32_t group_chunks(uint32_t g) { return (g == 0 || g == 4) ? 5u : (g == 3 || g == 7 || g == 8) ? 1u : 4u; }

struct PairParams {
    CUtensorMap map_obs, map_next;       // float16 rings as [n_worlds * capacity][160] tensors (TMA = true only)
    rl_world_cfg cfg;
    const int32_t* ev_rows;
    const int32_t* ev_total;
    rl_replay_bufs rp;
    const int32_t* sample_idx;
    rl_learn_bufs lb;
    const __half* wimg_e;
    const __half* wimg_t;
    long long* trace;      // debug: clock64 stamps of CTA 0, pair 3 (RL_TC_TRACE=1), else nullptr
};

__device__ __forceinline__ void sched_pair(int i, int& net, int& chunk) {
    if (i < 5) { net = 0; chunk = WI_W1 + i; }
    else if (i < 13) { net = 0; chunk = WI_W2N + (i - 5); }
    else if (i == 13) { net = 0; chunk = WI_WH; }
    else if (i < 19) { net = 1; chunk = WI_W1 + (i - 14); }
    else if (i < 27) { net = 1; chunk = WI_W2N + (i - 19); }
    else if (i == 27) { net = 1; chunk = WI_WH; }
    else if (i == 28) { net = 1; chunk = WI_WHT; }
    else { net = 1; chunk = WI_W2T + (i - 29); }
}

template <bool TMA>
__global__ void __launch_bounds__(TMA ? NTH_TMA : NTH, 1) k_learn_dueling_p(const __grid_constant__ PairParams P) {
    using L = Layout<RL_MODEL_DUELING>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // SWIZZLE_128B atoms are 1024-byte aligned
    __half* sH1 = reinterpret_cast<__half*>(smem + PO_H1);
    __half* sH2 = reinterpret_cast<__half*>(smem + PO_H2);
    __half* sStg = reinterpret_cast<__half*>(smem + PO_STG);
    __half* sDout = reinterpret_cast<__half*>(smem + PO_DOUT);
    float* bias_t = reinterpret_cast<float*>(smem + PO_BIAS);
    float* bias_e = bias_t + 400;
    float* red = reinterpret_cast<float*>(smem + PO_RED);
    int* meta = reinterpret_cast<int*>(smem + PO_META);
    unsigned long long* ringb = reinterpret_cast<unsigned long long*>(smem + PO_RING);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PO_BARS);
    // gfull[group]: all chunks of a release group have landed (one expect_tx per group: one mbarrier wait per stage on the issuer's
    // path); sfree[group]: every MMA that reads the chunks of that group has completed (one tcgen05.commit per GROUP of 4-5 chunks;
    // the 8-chunk stages are released in two halves so that the ring refills behind them)
    uint64_t* gfull = bars; uint64_t* sfree = bars + NSTG; uint64_t* done = sfree + NSTG; uint64_t* doneL1 = done + 1; uint64_t* go = done + 2;
    // xfull / xpfull: the gathered eval / target rows of a pair have landed; h2free / xfree: the MMAs that read the H2 / X region
    // have completed (tcgen05.commit): the next pair's rows may be gathered into it; metaready: the ring positions of a pair are in `meta`
    uint64_t* xfull = done + 3; uint64_t* xpfull = done + 4; uint64_t* h2free = done + 5; uint64_t* xfree = done + 6; uint64_t* metaready = done + 7;
    // w1a / w1b: the accumulator warps have taken columns 128..159 / 0..127 of the dW1 accumulator out of TMEM (gate the next target L1 / L2)
    uint64_t* w1a = done + 8; uint64_t* w1b = done + 9;
    // done2: the SECOND of two stages handed over back to back without a `go` in between (L2 half 1, the n2 >= 128 half of dWh).  A
    // parity-tracked mbarrier must never complete two phases before its waiters have seen the first one (they would wait for a flip
    // that has already happened twice), so consecutive completions alternate between `done` and `done2`.
    uint64_t* done2 = done + 10;
    // go2: H2[:, :128] of a forward pass is stored (epilogue of L2 half 0) -- the head's k-steps over those columns are issued behind
    // the MMAs of L2 half 1 instead of after the whole L2 epilogue.  Its own barrier for the same reason as done2: the next `go`
    // (L2 half 1 stored) may complete before the issuer has looked at this one.
    uint64_t* go2 = done + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* G = P.lb.grad_scratch + (size_t)blockIdx.x * L::N_TRAIN;
    const int total = *P.ev_total;
    const int n_my = total > (int)blockIdx.x ? (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_pairs = (n_my + 1) >> 1;
    const long long t_cta0 = clock64();

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTG; ++i) mbar_init(&gfull[i], 1);
        for (int i = 0; i < NSTG; ++i) mbar_init(&sfree[i], 1);
        mbar_init(done, 1); mbar_init(done2, 1); mbar_init(doneL1, 1); mbar_init(go, NEPI); mbar_init(go2, NEPI);
        mbar_init(xfull, NGI); mbar_init(xpfull, NGI); mbar_init(h2free, 1); mbar_init(xfree, 1); mbar_init(metaready, PB); mbar_init(w1a, 32 * NGW); mbar_init(w1b, 32 * NGW);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < NEPI) {
        reinterpret_cast<__half*>(smem + PO_ONES)[threadIdx.x] = __float2half_rn(1.0f);
        const float* Pe = P.lb.params; const float* Pt = P.lb.target;
        for (int i = threadIdx.x; i < 400; i += NEPI) {
            const int o = i < 128 ? L::OFF_B1 + i : i < 384 ? L::OFF_B2 + (i - 128) : L::OFF_BH + (i - 384);
            const bool ok = i < 384 + 9;
            bias_t[i] = ok ? Pt[o] : 0.f; bias_e[i] = ok ? Pe[o] : 0.f;
        }
        for (int i = threadIdx.x; i < L::N_TRAIN / 4; i += NEPI) reinterpret_cast<float4*>(G)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t aX = smem_u32(smem + PO_X), aH1 = smem_u32(sH1), aH2 = smem_u32(sH2), aD = smem_u32(sDout), aStg = smem_u32(sStg), aOnes = smem_u32(smem + PO_ONES);

    // (setmaxnreg is the first instruction of each role branch -- one per warpgroup, no control-flow merge behind it -- so that ptxas
    //  allocates each role's code against its own budget)
    if (warp >= 8 && warp < 12) {
      if (TMA) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_AUX));
      if (warp == 8) {
        // =================================== weight-stream producer ===================================
        if (lane == 0) {
            const uint32_t n_chunks = (uint32_t)n_pairs * SCHED_P;
            uint32_t freed = 0, stage = 0;                          // chunks of completed stages / next stage to wait for
            uint32_t grp = 0, left = 0;                             // release group of the chunk being produced / chunks left in it
            for (uint32_t produced = 0; produced < n_chunks; ++produced) {
                while (produced - freed >= (uint32_t)NSP) {
                    const uint32_t k = stage % NSTG;
                    mbar_wait(&sfree[k], (stage / NSTG) & 1);
                    freed += group_chunks(k);
                    ++stage;
                }
                if (left == 0) {                                    // first chunk of a group: arm its barrier with the group's bytes
                    left = group_chunks(grp);
                    fence_proxy_async();                            // earlier generic-proxy reads of the ring slots precede the async-proxy overwrite
                    mbar_expect_tx(&gfull[grp], left * (uint32_t)(HCH * 2));
                }
                int net, ch; sched_pair(produced % SCHED_P, net, ch);
                const uint32_t slot = produced % NSP;
                bulk_copy(sStg + slot * HCH, (net ? P.wimg_e : P.wimg_t) + (size_t)ch * HCH, HCH * 2, &gfull[grp]);
                if (--left == 0) grp = grp + 1 == NSTG ? 0 : grp + 1;
            }
        }
      } else if (warp == 9) {
        // =================================== MMA issuer (one warp, warp-uniform) ===================================
        // The whole warp walks the stage sequence; every value an MMA consumes is built from uniform values only and the MMA
        // itself is guarded by elect.sync, so descriptors stay in uniform registers and UTCHMMA issues without a convergence loop.
        // TMEM columns of the work area: L1 accumulators 128..255, L2 0..255, heads 0..15, dWh 0..31, dH2 half 0 128..255 /
        // half 1 0..127, dH1 0..127, dW1 0..159 -- the L1 of the eval net runs under the target head epilogue.
        const bool me = elect_one();
        const uint32_t T0 = __shfl_sync(0xffffffffu, *tmem_slot, 0), T_DW2 = T0 + 256;
        const int np_u = __shfl_sync(0xffffffffu, n_pairs, 0);
        uint32_t consumed = 0, go_no = 0, go2_no = 0, stage = 0, wstage = 0;
        int itr_n = 0, itr_p = -1;
        auto istamp = [&]() { if (P.trace && blockIdx.x == 0 && me && itr_p == 3 && itr_n < 40) P.trace[64 + itr_n++] = clock64(); };
        auto wait_go = [&]() { mbar_wait(go, go_no & 1); ++go_no; fence_after(); istamp(); };
        auto wait_go2 = [&]() { mbar_wait(go2, go2_no & 1); ++go2_no; fence_after(); };
        // all `n` chunks of the next release group have landed (one barrier per group; they are normally prefetched long before).
        // The wait for a stage's first group is placed BEFORE the stage's wait_go, off the hand-over path.
        auto chunks_wait = [&](int n) -> uint32_t {
            const uint32_t first = consumed;
            mbar_wait(&gfull[wstage % NSTG], (wstage / NSTG) & 1);
            ++wstage;
            consumed += n;
            istamp();
            return first;
        };
        auto chunk_addr = [&](uint32_t k) -> uint32_t { return aStg + (k % NSP) * (HCH * 2); };
        auto commit = [&](uint64_t* bar) { if (me) mma_commit(bar); };
        auto stage_free = [&]() { istamp(); commit(&sfree[stage % NSTG]); ++stage; };
        // L1: D[128 b][128 k1] = X[128][160] W1[128][160]^T -- 5 chunks [128 n][32 k], 2 k-steps each; X is a SWIZZLE_128B image
        auto l1 = [&](uint32_t ax, uint32_t k0) {
            const uint32_t id = idesc_h(128, 128, 0, 0);
#pragma unroll 1
            for (int c = 0; c < 5; ++c) {
                const uint64_t b = dk(chunk_addr(k0 + c), 32);
                if (me) {
                    mma_h(T0 + 128, dxk(ax, 2 * c), b, id, c != 0);
                    mma_h(T0 + 128, dxk(ax, 2 * c + 1), b + 16u, id, 1u);
                }
            }
            stage_free();
            commit(doneL1);
        };
        // L2: D[128 b][256 n2] = H1[128][128] W2[256][128]^T in two output halves (4 chunks [128 n][32 k] each, 2 k-steps per chunk):
        // half 0 -> columns 0..127, half 1 -> columns 128..255, each handed over on its own -- the epilogue of half 0 runs under the
        // MMAs of half 1
        auto l2 = [&](uint32_t k0) {
            const uint32_t id = idesc_h(128, 128, 0, 0);
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                if (hf) k0 = chunks_wait(4);
                uint64_t a = dk(aH1, 128);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const uint64_t b = dk(chunk_addr(k0 + c), 32);
                    if (me) { mma_h(T0 + 128 * hf, a, b, id, c != 0); mma_h(T0 + 128 * hf, a + 16u, b + 16u, id, 1u); }
                    a += 32u;
                }
                stage_free();
                commit(hf ? done2 : done);
            }
        };
        // head: D[128 b][16] = H2[128][256] Wh[16][256]^T -- one chunk, 16 k-steps in two halves: k-steps 0..7 read H2[:, :128] and are
        // issued once the epilogue of L2 half 0 has stored those columns (go2; they queue behind the MMAs of L2 half 1, whose accumulator
        // columns 128..255 they do not touch -- columns 0..15 have been drained by that epilogue), k-steps 8..15 after the whole L2 epilogue
        auto head = [&](uint32_t k0) {
            const uint32_t id = idesc_h(128, 16, 0, 0);
            uint64_t a = dk(aH2, 256), b = dk(chunk_addr(k0), 256);
            wait_go2();
#pragma unroll 1
            for (int ks = 0; ks < 16; ks += 4) {
                if (ks == 8) wait_go();
                if (me) {
                    mma_h(T0, a, b, id, ks != 0); mma_h(T0, a + 16u, b + 16u, id, 1u);
                    mma_h(T0, a + 32u, b + 32u, id, 1u); mma_h(T0, a + 48u, b + 48u, id, 1u);
                }
                a += 64u; b += 64u;
            }
            stage_free();
            commit(done);
        };
        for (int p = 0; p < np_u; ++p) {
            itr_p = p;
            uint32_t k0 = chunks_wait(5);                      // (the chunks of a stage are waited for BEFORE its go: off the hand-over path)
            wait_go();
            if (TMA) { if (p > 0) mbar_wait(w1a, (p - 1) & 1); mbar_wait(xpfull, p & 1); fence_after(); }
            l1(aH2, k0);                                       // target net: X' sits in the H2 region
            k0 = chunks_wait(4); wait_go();
            if (TMA && p > 0) { mbar_wait(w1b, (p - 1) & 1); fence_after(); }
            l2(k0);
            k0 = chunks_wait(1); head(k0);                     // target head (waits go2 / go inside), then the eval L1 (runs under the target head epilogue)
            k0 = chunks_wait(5);
            if (TMA) { mbar_wait(xfull, p & 1); fence_after(); }
            l1(aX, k0);
            k0 = chunks_wait(4); wait_go(); l2(k0);
            k0 = chunks_wait(1); head(k0);
            k0 = chunks_wait(1);
            wait_go();
            {   // dH2 half 0 (n2 < 128) -> columns 128..255: A = dOut [128][16], B = Wh^T chunk rows 0..127 [256 n2][16 j], and the half
                // n2 < 128 of dWh (rows n2 in [128 m, +128) -> columns 16 m..: A = H2 read MN-major, B = dOut read MN-major, K = 128 rows) are
                // handed over together: the dH2 epilogue overwrites H2[:, :128], which that half of dWh reads.  The half n2 >= 128 of dWh
                // runs under that epilogue.
                const uint32_t wht = chunk_addr(k0);
                const uint32_t id = idesc_h(128, 16, 1, 1);
                if (me) mma_h(T0 + 128, dk(aD, 16), dk(wht, 16), idesc_h(128, 128, 0, 0), 0u);
#pragma unroll 1
                for (int m = 0; m < 2; ++m) {
                    uint64_t a = dm(aH2 + m * 2048u, 256), b = dm(aD, 16);
#pragma unroll 1
                    for (int ks = 0; ks < 8; ks += 4) {
                        if (me) {
                            mma_h(T0 + 16 * m, a, b, id, ks != 0); mma_h(T0 + 16 * m, a + 512u, b + 32u, id, 1u);
                            mma_h(T0 + 16 * m, a + 1024u, b + 64u, id, 1u); mma_h(T0 + 16 * m, a + 1536u, b + 96u, id, 1u);
                        }
                        a += 2048u; b += 128u;
                    }
                    commit(m ? done2 : done);
                }
                wait_go();                                     // dH2[:, :128] stored, dWh drained: dH2 half 1 (n2 >= 128) -> columns 0..127
                if (me) mma_h(T0, dk(aD, 16), dk(wht + 4096u, 16), idesc_h(128, 128, 0, 0), 0u);
                stage_free();
                commit(done);
            }
            const uint32_t id_w2 = idesc_h(128, 128, 1, 1), id_h1 = idesc_h(128, 128, 0, 0);
            uint64_t a_h1 = dk(aH2, 256);
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                // The n2 half `hf` of dH2 is in place.  Under the epilogue of the other half (hf = 0) / at the end (hf = 1):
                // dW2[k1][n2 half] += sum_b H1[b][k1] dH2[b][n2] (both images MN-major, accumulator resident in TMEM) and the k-steps
                // n2 in [128 hf, +128) of dH1[128 b][128 k1] = dH2[128][256] W2^T[128 k1][256 n2]^T (4 chunks [128][32], 2 k-steps each,
                // accumulator columns 128..255: dH2 half 0 has been drained from them)
                if (hf) wait_go();
                uint64_t a2 = dm(aH1, 128), b2 = dm(aH2 + hf * 2048u, 256);
#pragma unroll 1
                for (int ks = 0; ks < 8; ks += 4) {
                    if (me) {
                        mma_h(T_DW2 + 128 * hf, a2, b2, id_w2, (p != 0 || ks != 0) ? 1u : 0u); mma_h(T_DW2 + 128 * hf, a2 + 256u, b2 + 512u, id_w2, 1u);
                        mma_h(T_DW2 + 128 * hf, a2 + 512u, b2 + 1024u, id_w2, 1u); mma_h(T_DW2 + 128 * hf, a2 + 768u, b2 + 1536u, id_w2, 1u);
                    }
                    a2 += 1024u; b2 += 2048u;
                }
                const uint32_t kc = chunks_wait(4);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const uint64_t b = dk(chunk_addr(kc + c), 32);
                    if (me) { mma_h(T0 + 128, a_h1, b, id_h1, (hf | c) != 0); mma_h(T0 + 128, a_h1 + 16u, b + 16u, id_h1, 1u); }
                    a_h1 += 32u;
                }
                stage_free();
            }
            commit(done);                                      // dW2 (reads H1) and dH1 are complete: the dH1 epilogue may overwrite H1
            {   // db2[n2] = sum_b dH2[b][n2] as a GEMM with a block of ones: A = dH2 read MN-major (M = n2 half m), B = ones [16 k][16]
                // (the same 512 bytes for every k-step) -> columns 16 m.. (all 16 equal); replaces 62 shuffles per 64 columns.  Its 16
                // issue-bound N = 16 MMAs run under the dH1 epilogue (handed over on done2)
                const uint32_t id = idesc_h(128, 16, 1, 1);
                const uint64_t b = dm(aOnes, 16);
#pragma unroll 1
                for (int m = 0; m < 2; ++m) {
                    uint64_t a = dm(aH2 + m * 2048u, 256);
#pragma unroll 1
                    for (int ks = 0; ks < 8; ks += 4) {
                        if (me) {
                            mma_h(T0 + 16 * m, a, b, id, ks != 0); mma_h(T0 + 16 * m, a + 512u, b, id, 1u);
                            mma_h(T0 + 16 * m, a + 1024u, b, id, 1u); mma_h(T0 + 16 * m, a + 1536u, b, id, 1u);
                        }
                        a += 2048u;
                    }
                }
                commit(done2);
                if (TMA) commit(h2free);                       // dH2 is dead: the next pair's target rows may land in the H2 region
            }
            wait_go();
            {   // dW1[k1][x] = sum_b dH1[b][k1] X[b][x]: dH1 (in the H1 region) read MN-major, X (SWIZZLE_128B) read MN-major, N = 160
                const uint32_t id = idesc_h(128, 160, 1, 1);
                uint64_t a = dm(aH1, 128);
#pragma unroll 1
                for (int ks = 0; ks < 8; ks += 4) {
                    if (me) {
                        mma_h(T0, a, dxm(aX, ks), id, ks != 0); mma_h(T0, a + 256u, dxm(aX, ks + 1), id, 1u);
                        mma_h(T0, a + 512u, dxm(aX, ks + 2), id, 1u); mma_h(T0, a + 768u, dxm(aX, ks + 3), id, 1u);
                    }
                    a += 1024u;
                }
                commit(done);
                if (TMA) commit(xfree);                        // X is dead: the next pair's eval rows may land in the X region
            }
        }
      } else if (TMA) {
        // =================================== warps 10, 11: row gatherers only ===================================
        // They take a sixth each of the 96 gathers of an image (gather op o = 6 l + j of lane l < 16, j = warp's gather index): the
        // accumulator warps then need ~1.6 k instead of ~2.4 k cycles to issue X' before they turn to the dW1 accumulator (w1a gates
        // the next target L1).
        const int j = NGW + (warp - 10);
        const int o = lane * NGI + j, g = o / 3, kb = o - g * 3;
        const bool act = lane < 16;
        for (int p = 0; p < n_pairs; ++p) {
            const int buf = p & 1;
            mbar_wait(metaready, p & 1);
            const int* ids = meta + buf * 4 * PB;
            int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
            if (act) {
                const int rb = (int)ringb[buf * 2 + (g >> 4)];
                const int4 i4 = *reinterpret_cast<const int4*>(ids + 4 * g);
                r0 = rb + i4.x; r1 = rb + i4.y; r2 = rb + i4.z; r3 = rb + i4.w;
            }
            if (p > 0) mbar_wait(h2free, (p - 1) & 1);
            if (lane == 0) mbar_expect_tx(xpfull, XIMG / NGI);
            __syncwarp();
            if (act) tma::gather4(aH2 + kb * XBLK + g * 512, &P.map_next, smem_u32(xpfull), kb * 64, r0, r1, r2, r3);
            if (p > 0) mbar_wait(xfree, (p - 1) & 1);
            if (lane == 0) mbar_expect_tx(xfull, XIMG / NGI);
            __syncwarp();
            if (act) tma::gather4(aX + kb * XBLK + g * 512, &P.map_obs, smem_u32(xfull), kb * 64, r0, r1, r2, r3);
        }
      }
    } else if (warp >= 12) {
        if (TMA) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_ACC));
        // =================================== dW1 accumulators + row gatherers (TMA = true), NGW = 4 warps ===================================
        // (a) dW1 lives in REGISTERS across the pairs of the CTA.  Flushing it with reds cost 80 KB of L2 atomics per pair (830 MB per
        //     launch, ~4 k cycles of LSU time per pair that also slowed the MMA stages running beside them).  Warp 12 + i covers the
        //     TMEM lane quarter i (= its warp id % 4): as soon as the dW1 MMAs of a pair have completed (xfree) each thread adds the 160
        //     accumulator columns of its row k1 to its registers -- columns 128..159 first (w1a: the next target L1 writes TMEM columns
        //     128..255), then 0..127 (w1b: the next target L2 writes 0..255); column 159 carries db1.  One plain store per element when
        //     the CTA is done.
        // (b) Row gathers: an image is 32 row groups (4 rows each) x three K blocks = 96 tile::gather4 (columns 160-191 are out of
        //     bounds and land as zeros).  A gather4 costs its issuing warp ~100 cycles and the rows land ~350 cycles after the last one
        //     is issued, so the 96 are spread over the NGW warps: warp i takes the row groups g = i (mod 4), lanes 0..23 of it one
        //     (row group, K block) each.  X' of pair p goes to the H2 region once pair p-1's dW2 / dH1 / db2 have completed (h2free),
        //     X of pair p to the X region once pair p-1's dW1 has completed (xfree).
        if (TMA) {
            const int w = warp - 12;
            const int o = lane * NGI + w, g = o / 3, kb = o - g * 3;       // gather op o of 96: row group g (4 rows), K block kb
            const bool act = lane < 16;
            if (lane == 0) tma::prefetch_map(w & 1 ? &P.map_obs : &P.map_next);
            const uint32_t T0 = *tmem_slot;
            const int k1 = (warp & 3) * 32 + lane;
            const uint32_t t_acc = T0 + ((uint32_t)((warp & 3) * 32) << 16);
            float acc[160];
#pragma unroll
            for (int j = 0; j < 160; ++j) acc[j] = 0.f;
            auto w1_add = [&]() {
#pragma unroll
                for (int i = 0; i < 10; ++i) {
                    const int cb = i < 2 ? 8 + i : i - 2;              // columns 128..159 first
                    float v[16];
                    tmem_ld16(t_acc + cb * 16, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[cb * 16 + j] += v[j];
                    if (i == 1) { fence_before(); mbar_arrive(w1a); }
                }
                fence_before();
                mbar_arrive(w1b);
            };
            for (int p = 0; p < n_pairs; ++p) {
                const int buf = p & 1;
                mbar_wait(metaready, p & 1);
                const int* ids = meta + buf * 4 * PB;
                int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
                if (act) {
                    const int rb = (int)ringb[buf * 2 + (g >> 4)];
                    const int4 i4 = *reinterpret_cast<const int4*>(ids + 4 * g);
                    r0 = rb + i4.x; r1 = rb + i4.y; r2 = rb + i4.z; r3 = rb + i4.w;
                }
                if (p > 0) mbar_wait(h2free, (p - 1) & 1);
                if (lane == 0) mbar_expect_tx(xpfull, XIMG / NGI);
                __syncwarp();
                if (act) tma::gather4(aH2 + kb * XBLK + g * 512, &P.map_next, smem_u32(xpfull), kb * 64, r0, r1, r2, r3);
                if (p > 0) { mbar_wait(xfree, (p - 1) & 1); fence_after(); w1_add(); }
                if (lane == 0) mbar_expect_tx(xfull, XIMG / NGI);
                __syncwarp();
                if (act) tma::gather4(aX + kb * XBLK + g * 512, &P.map_obs, smem_u32(xfull), kb * 64, r0, r1, r2, r3);
            }
            if (n_pairs > 0) {
                mbar_wait(xfree, (n_pairs - 1) & 1); fence_after();
                w1_add();
                G[L::OFF_B1 + k1] = acc[159] * (1.0f / H_SCALE); acc[159] = 0.f;                      // column 159 = db1
                float* gw = G + L::OFF_W1T + k1 * 4;                                                   // slab layout [x / 4][k1][x % 4]
#pragma unroll
                for (int j4 = 0; j4 < 40; ++j4)
                    *reinterpret_cast<float4*>(gw + j4 * 512) = make_float4(acc[j4 * 4] * (1.0f / H_SCALE), acc[j4 * 4 + 1] * (1.0f / H_SCALE),
                                                                             acc[j4 * 4 + 2] * (1.0f / H_SCALE), acc[j4 * 4 + 3] * (1.0f / H_SCALE));
            }
        }
    } else {
        // =================================== epilogue warps ===================================
        const uint32_t T0 = *tmem_slot, T_DW2 = T0 + 256;
        uint32_t done_no = 0, done2_no = 0, l1_no = 0;
        const int q = warp & 3, hh = warp >> 2;
        const int row = q * 32 + lane;                    // batch row of the pair == TMEM lane
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const int S = P.cfg.slot_cap, cap = P.rp.capacity;
        int tr_n = 0, tr_p = -1;
        auto stamp = [&]() { if (P.trace && blockIdx.x == 0 && threadIdx.x == 0 && tr_p == 3 && tr_n < 40) P.trace[tr_n++] = clock64(); };
        auto go_signal = [&]() { fence_proxy_async(); fence_before(); mbar_arrive(go); stamp(); };
        auto go2_signal = [&]() { fence_proxy_async(); fence_before(); mbar_arrive(go2); };
        auto wait_done = [&]() { mbar_wait(done, done_no & 1); ++done_no; fence_after(); stamp(); };
        auto wait_done2 = [&]() { mbar_wait(done2, done2_no & 1); ++done2_no; fence_after(); stamp(); };
        auto wait_l1 = [&]() { mbar_wait(doneL1, l1_no & 1); ++l1_no; fence_after(); stamp(); };
        // ---- metadata of a pair (threads 0-127, one sampled row each) in three phases so that no global-load latency
        //      sits on the critical path: A event row + ring position, B action / reward / done (+ L2 prefetch of the
        //      sample's two rows), C to shared memory (+ metaready for the row gatherer) ----
        size_t m_ring = 0; int m_row = 0, m_i = 0, m_act = 0; float m_rew = 0.f, m_dn = 0.f;
        auto meta_a = [&](int p) {
            if (threadIdx.x < PB) {
                const int r = threadIdx.x, ev = r >> 6;
                const int it = 2 * p + ev;
                const int e = (int)blockIdx.x + (it < n_my ? it : 2 * p) * (int)gridDim.x;     // a missing second event mirrors the first
                m_row = P.ev_rows[e];
                m_i = P.sample_idx[(size_t)e * R + (r & 63)];       // (no dependent instruction here: the loads stay in flight)
            }
        };
        auto meta_b = [&]() {
            if (threadIdx.x < PB) {
                m_ring = (size_t)(m_row / S) * cap; m_i = max(m_i, 0);
                m_act = P.rp.action[m_ring + m_i]; m_rew = P.rp.reward[m_ring + m_i]; m_dn = (float)P.rp.done[m_ring + m_i];
                // pull the two rows of this sample towards L2 now: the gathers of the next pair then hit L2, not HBM
                const size_t rb = (m_ring + m_i) * RL_K1 * (P.rp.obs_fp16 ? 2 : 4);                  // byte offset of the sample's row
                const char* r0 = reinterpret_cast<const char*>(P.rp.next_obs) + rb; const char* r1 = reinterpret_cast<const char*>(P.rp.obs) + rb;
                const int nl = P.rp.obs_fp16 ? 3 : 5;
                for (int ln = 0; ln < nl; ++ln) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + ln * 128));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + ln * 128));
                }
            }
        };
        auto meta_c = [&](int buf) {
            if (threadIdx.x < PB) {
                const int r = threadIdx.x;
                int* m = meta + buf * 4 * PB;
                m[r] = m_i; m[PB + r] = m_act;
                reinterpret_cast<float*>(m)[2 * PB + r] = m_rew; reinterpret_cast<float*>(m)[3 * PB + r] = m_dn;
                if ((r & 63) == 0) ringb[buf * 2 + (r >> 6)] = (unsigned long long)m_ring;
                if (TMA) mbar_arrive(metaready);
            }
        };
        // ---- register gather (float32 rings only, TMA = false): 128 rows x 20 units of 8 columns -> packed fp16; a quarter-warp
        //      takes 8 consecutive units of a row (contiguous loads, conflict-free swizzled stores) ----
        uint4 xh[TMA ? 1 : 10];
        auto gather_load = [&](const float* __restrict__ src, int buf) {
            if (TMA) return;
            const int* ids = meta + buf * 4 * PB;
            const bool ring16 = P.rp.obs_fp16 != 0;
#pragma unroll
            for (int u = 0; u < (TMA ? 0 : 10); ++u) {
                const int v = threadIdx.x + u * NEPI;
                const int r = v / 20, oct = v - r * 20;
                const size_t ro = (size_t)ringb[buf * 2 + (r >> 6)] + ids[r];
                if (ring16) {
                    xh[u] = __ldg(reinterpret_cast<const uint4*>(src) + ro * (RL_K1 / 8) + oct);
                } else {
                    const float4* g = reinterpret_cast<const float4*>(src + ro * RL_K1) + oct * 2;
                    const float4 a = __ldg(g), b = __ldg(g + 1);
                    xh[u] = make_uint4(pk(a.x, a.y), pk(a.z, a.w), pk(b.x, b.y), oct == 19 ? pk(b.z, 1.0f) : pk(b.z, b.w));   // column 159 := 1 (db1)
                }
            }
        };
        auto gather_store = [&](int region) {
            if (TMA) return;
#pragma unroll
            for (int u = 0; u < (TMA ? 0 : 10); ++u) {
                const int v = threadIdx.x + u * NEPI;
                const int r = v / 20, oct = v - r * 20;
                *reinterpret_cast<uint4*>(smem + region + ximg(r, oct)) = xh[u];
            }
        };
        // relu(D + bias) of 32 accumulator columns -> 4 x 16-byte stores into a batch-major image of width K
        auto relu_store32 = [&](float (&v)[32], const float* bias, __half* img, int c0, int K) {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                const float4 b0 = *reinterpret_cast<const float4*>(bias + j8 * 8), b1 = *reinterpret_cast<const float4*>(bias + j8 * 8 + 4);
                *reinterpret_cast<uint4*>(img + himg(row, c0 + j8 * 8, K)) =
                    make_uint4(pk_relu(v[j8 * 8] + b0.x, v[j8 * 8 + 1] + b0.y), pk_relu(v[j8 * 8 + 2] + b0.z, v[j8 * 8 + 3] + b0.w),
                               pk_relu(v[j8 * 8 + 4] + b1.x, v[j8 * 8 + 5] + b1.y), pk_relu(v[j8 * 8 + 6] + b1.z, v[j8 * 8 + 7] + b1.w));
            }
        };
        // ---- L1 epilogue: this thread's row, columns [64 hh, 64 hh + 64) of the accumulator at TMEM columns 128..255 ----
        auto l1_epilogue = [&](const float* bias) {
            const int c0 = hh * 64;
            float v0[32], v1[32];
            tmem_ld32(T0 + t_lane + 128 + c0, v0);
            tmem_ld32(T0 + t_lane + 128 + c0 + 32, v1);
            tmem_wait_ld();
            relu_store32(v0, bias + c0, sH1, c0, 128);
            relu_store32(v1, bias + c0 + 32, sH1, c0 + 32, 128);
        };
        // ---- L2 epilogue, one call per output half (each handed over on its own): this thread's row, columns
        //      [128 half + 64 hh, +64): H2 = relu(D + b2) ----
        auto l2_epilogue = [&](const float* bias, int half) {
            const int c0 = half * 128 + hh * 64;
            float v0[32], v1[32];
            tmem_ld32(T0 + t_lane + c0, v0);
            tmem_ld32(T0 + t_lane + c0 + 32, v1);
            tmem_wait_ld();
            relu_store32(v0, bias + 128 + c0, sH2, c0, 256);
            relu_store32(v1, bias + 128 + c0 + 32, sH2, c0 + 32, 256);
        };
        // ---- head epilogue (warps 0-3: one row each): out[0..8] of the row, mean of the advantages over the row's EVENT
        //      (PERD3QN.py:202: advantage.mean() over the whole [64, 8] tensor) ----
        auto head_epilogue = [&](const float* bias, float (&out)[9]) -> float {
            float v[16];
            tmem_ld16(T0 + t_lane, v);
            tmem_wait_ld();
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 9; ++j) { out[j] = v[j] + bias[384 + j]; if (j < 8) s += out[j]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[q] = s;
            head_bar();
            const float m = (q < 2 ? red[0] + red[1] : red[2] + red[3]) * (1.0f / (8 * R));
            head_bar();
            return m;
        };

        // gradients whose owner (thread, element) is the same in every pair are summed in registers and leave once per CTA:
        // dWh row n2 = 128 hh + row (9), db2[n2] of the same row (column 0 of the ones GEMM), dbh[lane], db1[row]
        float acc_wh[9], acc_b2 = 0.f, acc_bh = 0.f, acc_b1 = 0.f;
#pragma unroll
        for (int j = 0; j < 9; ++j) acc_wh[j] = 0.f;
        if (n_pairs > 0) {
            meta_a(0); meta_b(); meta_c(0);
            epi_bar();
            gather_load(P.rp.next_obs, 0);
            gather_store(PO_H2);
            go_signal();                                            // -> target L1 of pair 0
        }
        for (int p = 0; p < n_pairs; ++p) {
            const int buf = p & 1;
            tr_p = p; stamp();
            const int* idx = meta + buf * 4 * PB; const int* act = idx + PB;
            const float* rew = reinterpret_cast<const float*>(idx + 2 * PB); const float* dn = rew + PB;
            const bool more = p + 1 < n_pairs;
            const int ev = row >> 6;
            const bool valid = 2 * p + ev < n_my;
            const int e = (int)blockIdx.x + (2 * p + ev) * (int)gridDim.x;
            // ---------------- target net ----------------
            if (more) meta_a(p + 1);
            wait_l1();
            l1_epilogue(bias_t);
            go_signal();                                            // -> target L2
            if (more) meta_b();
            wait_done();
            gather_load(P.rp.obs, buf);                             // (float32 rings: eval rows of this pair, in flight behind the L2 epilogue)
            l2_epilogue(bias_t, 0);
            go2_signal();                                           // -> first half of the target head's k-steps
            wait_done2();
            l2_epilogue(bias_t, 1);
            gather_store(PO_X);
            go_signal();                                            // -> target head, eval L1
            if (more) meta_c(buf ^ 1);
            wait_done();
            float nq = 0.f;
            if (hh == 0) {
                float o[9];
                const float mean = head_epilogue(bias_t, o);
                float mx = o[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) mx = fmaxf(mx, o[j]);
                nq = mx + o[8] - mean;
            }
            // ---------------- eval net ----------------
            wait_l1();
            l1_epilogue(bias_e);
            go_signal();                                            // -> eval L2 (overwrites the target head columns: consumed above)
            wait_done();
            l2_epilogue(bias_e, 0);
            go2_signal();                                           // -> first half of the eval head's k-steps
            wait_done2();
            l2_epilogue(bias_e, 1);
            go_signal();                                            // -> eval head
            wait_done();
            float dbh = 0.f;                                        // lane j < 9 of warps 0-3: sum over the warp's rows of dOut[.][j]
            if (hh == 0) {
                // ---- TD target, loss, priorities, dOut of this row (PERD3QN.py:103-110), all thread-local ----
                float o[9];
                const float mean = head_epilogue(bias_e, o);
                const int a = act[row];
                float qsel = o[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) qsel = (a == j) ? o[j] : qsel;
                const float qa = qsel + o[8] - mean;
                const float y = rew[row] + P.lb.gamma * (1.0f - dn[row]) * nq;
                const float diff = qa - y;
                const float g = valid ? 2.0f * diff * (1.0f / R) : 0.f;
                float sq = diff * diff, gs = g;
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) { sq += __shfl_xor_sync(0xffffffffu, sq, o2); gs += __shfl_xor_sync(0xffffffffu, gs, o2); }
                if (lane == 0) { red[8 + q] = sq; red[16 + q] = gs; }
                head_bar();
                const float loss = (q < 2 ? red[8] + red[9] : red[10] + red[11]) * (1.0f / R);
                const float shift = valid ? (q < 2 ? red[16] + red[17] : red[18] + red[19]) * (1.0f / (8 * R)) : 0.f;
                if (valid) {
                    P.lb.new_prio[(size_t)e * R + (row & 63)] = fabsf(nq - qa);
                    if ((row & 63) == 0) P.lb.loss[e] = loss;
                }
                float d[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) d[j] = j < 8 ? ((j == a ? g : 0.f) - shift) : (j == 8 ? g : 0.f);
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = pk_sat(d[2 * j] * H_SCALE, d[2 * j + 1] * H_SCALE);
                *reinterpret_cast<uint4*>(sDout + himg(row, 0, 16)) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(sDout + himg(row, 8, 16)) = make_uint4(w[4], w[5], w[6], w[7]);
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    float s = d[j];
#pragma unroll
                    for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
                    dbh = lane == j ? s : dbh;
                }
            }
            go_signal();                                            // -> dH2 half 0, dWh
            acc_bh += dbh;
            // ---- dH2 epilogue, two halves of 128 features: dH2 = H2 > 0 ? acc : 0 (packed: F2FP.SATFINITE, HSET2, HMUL2), in place of H2 ----
            auto dh2_epilogue = [&](uint32_t tcol, int n0) {           // this thread's row, 64 columns: accumulator tcol.., features n0..
                float va[32], vb[32];
                tmem_ld32(T0 + t_lane + tcol, va);
                tmem_ld32(T0 + t_lane + tcol + 32, vb);
                uint4 hm[8];
#pragma unroll
                for (int j8 = 0; j8 < 8; ++j8) hm[j8] = *reinterpret_cast<const uint4*>(sH2 + himg(row, n0 + j8 * 8, 256));
                tmem_wait_ld();
#pragma unroll
                for (int j8 = 0; j8 < 8; ++j8) {
                    const float* v = j8 < 4 ? va + j8 * 8 : vb + (j8 - 4) * 8;
                    *reinterpret_cast<uint4*>(sH2 + himg(row, n0 + j8 * 8, 256)) =
                        make_uint4(mask_pos(pk_sat(v[0], v[1]), hm[j8].x), mask_pos(pk_sat(v[2], v[3]), hm[j8].y),
                                   mask_pos(pk_sat(v[4], v[5]), hm[j8].z), mask_pos(pk_sat(v[6], v[7]), hm[j8].w));
                }
            };
            wait_done();                                            // dH2 half 0
            dh2_epilogue(128u + hh * 64, hh * 64);
            wait_done2();                                           // the n2 >= 128 half of dWh (ran under the epilogue above)
            float wv[16];
            tmem_ld16(T0 + t_lane + 16 * hh, wv);                   // head weight gradients: row n2 = 128 hh + this thread's TMEM lane
            tmem_wait_ld();
            go_signal();                                            // -> dH2 half 1 (its accumulator overwrites the dWh columns); dW2 / dH1 over n2 < 128
#pragma unroll
            for (int j = 0; j < 9; ++j) acc_wh[j] += wv[j];
            wait_done();                                            // dH2 half 1
            dh2_epilogue(hh * 64, 128 + hh * 64);
            go_signal();                                            // -> dW2 / dH1 over n2 >= 128, db2
            if (!TMA && more) epi_bar();                            // the next pair's metadata (threads 0-127) is visible to every warp
            wait_done();
            if (more) gather_load(P.rp.next_obs, buf ^ 1);          // (float32 rings: next pair's target rows, in flight behind the dH1 epilogue)
            {   // the dH1 epilogue: this thread's row, columns [64 hh, +64) of the accumulator at TMEM columns 128..255:
                // dH1 = H1 > 0 ? acc : 0 in place of H1
                const int c0 = hh * 64;
                float v0[32], v1[32];
                tmem_ld32(T0 + t_lane + 128 + c0, v0);
                tmem_ld32(T0 + t_lane + 128 + c0 + 32, v1);
                uint4 hm[8];
#pragma unroll
                for (int j8 = 0; j8 < 8; ++j8) hm[j8] = *reinterpret_cast<const uint4*>(sH1 + himg(row, c0 + j8 * 8, 128));
                tmem_wait_ld();
#pragma unroll
                for (int j8 = 0; j8 < 8; ++j8) {
                    const float* v = j8 < 4 ? v0 + j8 * 8 : v1 + (j8 - 4) * 8;
                    *reinterpret_cast<uint4*>(sH1 + himg(row, c0 + j8 * 8, 128)) =
                        make_uint4(mask_pos(pk_sat(v[0], v[1]), hm[j8].x), mask_pos(pk_sat(v[2], v[3]), hm[j8].y),
                                   mask_pos(pk_sat(v[4], v[5]), hm[j8].z), mask_pos(pk_sat(v[6], v[7]), hm[j8].w));
                }
            }
            wait_done2();                                           // db2 (ran under the epilogue above): feature n2 = 128 hh + row, before dW1 overwrites its columns
            {
                float b2v[16];
                tmem_ld16(T0 + t_lane + 16 * hh, b2v);
                tmem_wait_ld();
                acc_b2 += b2v[0];
            }
            if (more) gather_store(PO_H2);                          // (float32 rings) X' of the next pair -> H2 region: dH2 is dead (dW2 / dH1 / db2 done)
            go_signal();                                            // -> dW1
            wait_done();
            if (TMA) {   // the accumulator warps take dW1 out of TMEM (w1a, w1b gate the next pair's target L1, L2)
                if (more) go_signal();
            } else {     // float32 rings (no gather warps): columns [80 hh, +80); column 159 carries db1.  All 80 values are pulled into
                // registers first so that the next pair's target L1 can start before the reds are issued.
                float v[80];
#pragma unroll
                for (int cb = 0; cb < 5; ++cb) tmem_ld16(T0 + t_lane + hh * 80 + cb * 16, v + cb * 16);
                tmem_wait_ld();
                if (more) go_signal();                              // -> next pair's target L1
                // slab layout [x / 4][k1][x % 4]: the 32 lanes of a red.v4 cover 512 contiguous bytes (16 sectors instead of 32)
                float* gw = G + L::OFF_W1T + ((hh * 20) * 128 + row) * 4;
#pragma unroll
                for (int j = 0; j < 80; ++j) v[j] *= (1.0f / H_SCALE);
                if (hh == 1) { acc_b1 += v[79]; v[79] = 0.f; }                  // (already unscaled)
#pragma unroll
                for (int j4 = 0; j4 < 20; ++j4) red_add4(gw + j4 * 512, v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            }
        }
        if (n_pairs > 0) {
            // (global reds are kept off the per-pair path: an mbarrier arrive has release semantics and waits for the thread's
            //  outstanding reds, and scattered reds share the LSU with the epilogue's shared-memory traffic)
            const int n2 = hh * 128 + row;
#pragma unroll
            for (int j = 0; j < 9; ++j) G[L::OFF_WH + n2 * 9 + j] = acc_wh[j] * (1.0f / H_SCALE);          // sole owner of the row
            G[L::OFF_B2 + n2] = acc_b2 * (1.0f / H_SCALE);                                               // sole owner
            if (hh == 0 && lane < 9) red_add(G + L::OFF_BH + lane, acc_bh);
            if (!TMA && hh == 1) G[L::OFF_B1 + row] = acc_b1;      // (TMA: the accumulator warps own db1)
        }
        if (n_pairs > 0) {     // flush the TMEM-resident dW2 accumulator once: lane = k1, columns [128 hh, +128) of n2
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
                const int c0 = hh * 128 + cb * 32;
                float v[32];
                tmem_ld32(T_DW2 + t_lane + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    *reinterpret_cast<float4*>(G + L::OFF_W2T + row * 256 + c0 + j4 * 4) =
                        make_float4(v[j4 * 4] * (1.0f / H_SCALE), v[j4 * 4 + 1] * (1.0f / H_SCALE), v[j4 * 4 + 2] * (1.0f / H_SCALE), v[j4 * 4 + 3] * (1.0f / H_SCALE));
            }
        }
    }
    fence_before();
    __syncthreads();
    if (P.trace && threadIdx.x == 0 && blockIdx.x < 160) { P.trace[128 + 2 * blockIdx.x] = clock64() - t_cta0; P.trace[129 + 2 * blockIdx.x] = n_pairs; }
    if (warp == 8) tmem_dealloc(*tmem_slot, 512);
}

}  // namespace

extern "C" int rl_brain_learn_p(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                                const int32_t* sample_idx, const rl_learn_bufs* learn, const void* wimg_eval_h,
                                const void* wimg_target_h, void* stream) {
    RL_ARG_CHECK(cfg && rows && replay && sample_idx && learn && wimg_eval_h && wimg_target_h);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1 && learn->batch == R);
    RL_ARG_CHECK(learn->params && learn->target && learn->grad_scratch && learn->grad && learn->new_prio && learn->loss);
    if (learn->kind != RL_MODEL_DUELING) return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_learn_p: dueling networks only");
    PairParams P;
    memset(&P, 0, sizeof(P));
    P.cfg = *cfg;
    P.ev_rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_EVENT) * rows->row_cap;
    P.ev_total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    P.rp = *replay; P.sample_idx = sample_idx; P.lb = *learn;
    P.wimg_e = reinterpret_cast<const __half*>(wimg_eval_h); P.wimg_t = reinterpret_cast<const __half*>(wimg_target_h);
    P.trace = nullptr;
    const bool use_tma = replay->obs_fp16 != 0 && getenv("RL_PAIR_NO_TMA") == nullptr;     // (A/B switch: register gather from the float16 ring)
    if (use_tma) {
        const long long n_rows = (long long)cfg->n_worlds * replay->capacity;
        RL_ARG_CHECK(n_rows > 0 && n_rows < (1ll << 31));
        if (tma::make_rows_map(&P.map_obs, replay->obs, (uint64_t)n_rows, RL_K1, 1) != 0 ||
            tma::make_rows_map(&P.map_next, replay->next_obs, (uint64_t)n_rows, RL_K1, 1) != 0)
            return rl_set_err(RL_ERR_CUDA, "rl_brain_learn_p: cuTensorMapEncodeTiled failed");
    }
    static long long* trace_dev = nullptr;
    const bool tracing = getenv("RL_TC_TRACE") != nullptr;
    if (tracing) {
        if (!trace_dev) RL_CUDA_CHECK(cudaMalloc(&trace_dev, 512 * sizeof(long long)));
        RL_CUDA_CHECK(cudaMemset(trace_dev, 0, 512 * sizeof(long long)));
        P.trace = trace_dev;
    }
    static PerDeviceOnce attr;
    if (attr.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_learn_dueling_p<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_learn_dueling_p<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM));
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (use_tma) k_learn_dueling_p<true><<<rl_learn_grid(), NTH_TMA, PAIR_SMEM, st>>>(P);
    else k_learn_dueling_p<false><<<rl_learn_grid(), NTH, PAIR_SMEM, st>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    if (tracing) {
        long long h[512];
        RL_CUDA_CHECK(cudaStreamSynchronize(st));
        RL_CUDA_CHECK(cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[pair trace] epilogue stamps (cycles since the pair's start):");
        for (int i = 1; i < 40 && h[i]; ++i) fprintf(stderr, " %lld", h[i] - h[0]);
        fprintf(stderr, "\n[pair trace] issuer stamps (go seen / chunks landed / last MMA of a chunk stage issued):");
        for (int i = 0; i < 40 && h[64 + i]; ++i) fprintf(stderr, " %lld", h[64 + i] - h[0]);
        long long cmin = 1ll << 60, cmax = 0, csum = 0; int nc = 0;
        for (int c = 0; c < 160; ++c) if (h[128 + 2 * c]) { const long long v = h[128 + 2 * c]; cmin = v < cmin ? v : cmin; cmax = v > cmax ? v : cmax; csum += v; ++nc; }
        fprintf(stderr, "\n[pair trace] cycles per CTA: min %lld avg %lld max %lld over %d CTAs; CTA 0: %lld cycles, %lld pairs\n", cmin, nc ? csum / nc : 0, cmax, nc,
                h[128], h[129]);
    }
    return rl_learn_reduce(learn, P.ev_total, 2, (void*)st);
}

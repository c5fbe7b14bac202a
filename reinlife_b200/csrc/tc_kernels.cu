// tc_kernels.cu -- the dueling brains (PERD3QN / D3QN) on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators,
// cp.async.bulk weight streaming), same math and outputs as the fp32 kernels of learn_kernels.cu / brain_kernels.cu, which
// stay as the tight-tolerance reference path:
//   k_learn_dueling_h     train() events, fp16 operands (kind::f16)      -- the default event kernel
//   k_learn_dueling_tc2   train() events, tf32 operands (kind::tf32)     -- precision="tf32"
//   k_act_dueling_tc      get_action forward, tf32 operands
//   k_build_wimg_dueling(_h)  weight operand images (23 chunks per network) from the kernel-layout parameters
// One event (64 rows) per CTA iteration.  The large GEMMs of an event -- target/eval L1 160->128, L2 128->256, head 256->9,
// dH2 = dOut Wh^T, dH1 = dH2 W2, dW2 += H1^T dH2 (accumulated in TMEM across ALL events of the CTA), dW1^T = dH1^T X and,
// in the fp16 kernel, dWh = H2^T dOut -- run as K-major MMAs in transposed-output form (see k_learn_dueling_tc2).
// tf32 MN-major operands read back as zeros on this part with the no-swizzle layouts (tests/test_tc_gpu.py pins the
// K-major conventions), so transposed operand images are written explicitly by the epilogues.
#include <stdlib.h>
#include <cuda_fp16.h>
#include "tc_tile.cuh"
#include "models.cuh"

namespace {

using namespace tc;
using mlp::mbar_init; using mlp::mbar_wait; using mlp::fence_mbar_init; using mlp::fence_proxy_async; using mlp::bulk_load;

constexpr int R = 64;
constexpr int NEPI = 256;                // epilogue threads (warps 0-7)
constexpr int NTHREADS = NEPI + 32;      // + producer warp
constexpr int NS = 4;                    // weight-chunk stages
constexpr int CHUNK_F = 4096;            // floats per chunk (16 KB)

// ---- weight image buffer (per network), in chunks of 16 KB ----
constexpr int WI_W1 = 0;     // 5 chunks [128 n][32 k]   B of L1:   n = k1, k = kx
constexpr int WI_W2K = 5;    // 8 chunks [256 n][16 k]   B of L2:   n = n2, k = k1
constexpr int WI_WH = 13;    // 1 chunk  [16 n][256 k]   B of head: n = j,  k = n2
constexpr int WI_WHT = 14;   // 1 chunk  [256 n][16 k]   B of dH2:  n = n2, k = j
constexpr int WI_W2T = 15;   // 8 chunks [128 n][32 k]   B of dH1:  n = k1, k = n2
constexpr int WI_CHUNKS = 23;
constexpr int WI_W2N = 23;   // fp16 images only: 8 chunks [128 n][32 k]  B of L2 split by output halves: chunk 4 h + q = n2 in [128 h, +128), k1 in [32 q, +32)
constexpr int WI_CHUNKS_H = 31;

// chunk schedule of one event: (net, chunk).  net 0 = target, 1 = eval
constexpr int SCHED_N = 37;

struct TcLearnParams {
    rl_world_cfg cfg;
    const int32_t* ev_rows;
    const int32_t* ev_total;
    rl_replay_bufs rp;
    const int32_t* sample_idx;
    rl_learn_bufs lb;
    const float* wimg_e;
    const float* wimg_t;
    long long* trace;      // debug: clock64 stamps of CTA 0 (RL_TC_TRACE=1), else nullptr
};

// ---- shared memory carve-up (floats).  Region reuse over one event:
//   sX  : X' (target) -> X (eval) -> H1^T (eval, written by the L1 epilogue once the L1 MMAs are done) -> dH1^T
//   sH1 : H1 (target) -> H1 (eval) -> dH2^T half buffer [128][64]
//   sH2 : H2 (target) -> H2 (eval) -> dH2 -> X^T [160][64] (after the dH1 MMAs)
constexpr int SM_X = 0;                          // 64x160                                10240
constexpr int SM_H1 = SM_X + 10240;              // [64][128]                              9216
constexpr int SM_H2 = SM_H1 + 9216;              // [64][256]                             16384   (H1 region: 36 KB for the padded dH2^T half image)
constexpr int SM_STAGE = SM_H2 + 16384;          // NS x 4096
constexpr int SM_DOUT = SM_STAGE + NS * CHUNK_F; // dOut image [64][16] (K-major A)       1024
constexpr int SM_OUTH = SM_DOUT + 1024;          // head outputs [64][16]                 1024
constexpr int SM_DPL = SM_OUTH + 1024;           // dOut plain [64][12]                    768
constexpr int SM_SMALL = SM_DPL + 768;           // rew, dn, nq, gb [4][64] + red[32] + (b1[128] b2[256] bh[16]) x 2 nets
constexpr int SM_SMALL_N = 4 * 64 + 32 + 2 * (128 + 256 + 16);
constexpr int SM_INT = SM_SMALL + SM_SMALL_N;    // per-event metadata, double buffered: [2] x { idx[64], act[64], rew[64], dn[64] }
constexpr int SM_FLOATS = SM_INT + 2 * 256;
constexpr size_t TC_SMEM = sizeof(float) * SM_FLOATS + 8 * (2 * NS + 2) + 16;
static_assert(TC_SMEM <= 227 * 1024, "shared memory budget");

// fire-and-forget adds into the CTA-private gradient slab (no return value -> no scoreboard stall, no contention)
__device__ __forceinline__ void red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// Transposed operand images ([feature rows][64 batch columns]) use a padded chunk stride: LBO = 144 B instead of 128 B.
// A transposed store writes one feature row for 16 consecutive batch columns; with the pad the four 16-byte chunks land
// in different banks (conflict-free) -- the descriptor's LBO/SBO make the padding invisible to the tensor core.
constexpr int TP_CH = 36;                  // floats between consecutive 4-column chunks
constexpr int TP_RG = 16 * TP_CH;          // floats per 8-row group (64 columns = 16 chunks)
__device__ __forceinline__ int timg_off(int r, int c) { return (r >> 3) * TP_RG + (c >> 2) * TP_CH + (r & 7) * 4 + (c & 3); }
__device__ __forceinline__ uint64_t desc_timg(uint32_t saddr) { return make_desc(saddr, TP_CH * 4, TP_RG * 4); }
constexpr int TP_KSTEP = 2 * TP_CH * 4;    // bytes per MMA k-step (8 batch columns = 2 chunks)

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float epi_sum(float v, float* red) {     // sum over the 256 epilogue threads
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    epi_bar();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += red[i];
    epi_bar();
    return r;
}

// gather 64 rows of 160 floats into an interleaved image (TRANSPOSED = false: [64][160]; true: [160][64]).
// The rows come from the replay ring in HBM: all ten 16-byte loads of a thread are issued before the first store,
// so a gather costs one DRAM round trip instead of ten.
// The gather is split in two so that the DRAM/L2 round trip hides behind a tensor-core stage: gather_load issues the
// ten 16-byte loads of a thread (rows come from the replay ring in HBM), gather_store writes the image later.
__device__ __forceinline__ void gather_load(float4 (&x)[10], const float* __restrict__ src, const int* ids) {
#pragma unroll
    for (int u = 0; u < 10; ++u) {
        const int v = threadIdx.x + u * NEPI;
        const int rr = v & 7, cc = (v >> 3) & 3, blk = v >> 5;         // blk: 0..79 -> (row group 0..7, chunk group 0..9)
        const int rg = blk / 10, cg = blk - rg * 10;
        x[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)ids[rg * 8 + rr] * RL_K1) + cg * 4 + cc);
    }
}
template <bool TRANSPOSED>
__device__ __forceinline__ void gather_store(float* img, const float4 (&x)[10]) {
    // lane -> (row within an 8-row group, 4 consecutive 16-byte chunks): conflict-free stores for the plain image
#pragma unroll
    for (int u = 0; u < 10; ++u) {
        const int v = threadIdx.x + u * NEPI;
        const int rr = v & 7, cc = (v >> 3) & 3, blk = v >> 5;
        const int rg = blk / 10, cg = blk - rg * 10;
        const int r = rg * 8 + rr, c4 = cg * 4 + cc;
        if (!TRANSPOSED) {
            *reinterpret_cast<float4*>(img + img_off(r, c4 * 4, RL_K1)) = make_float4(to_tf32(x[u].x), to_tf32(x[u].y), to_tf32(x[u].z), to_tf32(x[u].w));
        } else {
            img[timg_off(c4 * 4 + 0, r)] = to_tf32(x[u].x); img[timg_off(c4 * 4 + 1, r)] = to_tf32(x[u].y);
            img[timg_off(c4 * 4 + 2, r)] = to_tf32(x[u].z); img[timg_off(c4 * 4 + 3, r)] = to_tf32(x[u].w);
        }
    }
}

// issue a tensor-core stage from epilogue thread 0 (after stage_sync), everybody then waits for its completion (used by the
// get_action kernel, whose stages are short and strictly sequential)
#define RL_STAGE(BODY) do { stage_sync(); if (threadIdx.x == 0) { fence_after(); BODY; mma_commit(done); } } while (0)

// =====================================================================================================
// k_learn_dueling_tc2 -- the same train() event with every activation GEMM in TRANSPOSED-OUTPUT form:
//   D[feature][batch] = W[feature][k] * Act[batch][k]^T   (A = weight chunk, M = 128 features; B = activation image, N = 64 rows)
// so that (1) every MMA runs at M = 128 (full tensor rate, M = 64 runs at half), (2) the TMEM epilogues own one FEATURE
// per lane -- all 32 lanes of all 8 warps carry data (an M = 64 accumulator layout fills 16 of 32),
// bias is a per-thread scalar, bias gradients are per-thread sums, (3) the feature-major images the weight-gradient GEMMs
// need (H1^T, dH2^T, dH1^T) are written with 16-byte vector stores, and the batch-major images the next layer needs
// (H1, H2, dH2 as B operands) by conflict-free scalar scatters into images with a padded chunk stride (LBO = 144 B).
// =====================================================================================================
constexpr int NS2 = 4;                                  // weight-chunk stages
constexpr int V2_X = 0;                                 // X' / X [64][160] plain image -> H1^T / dH1^T timg [128][64]   10240
constexpr int V2_H1 = V2_X + 10240;                     // H1 bimg [64][128] -> dH2^T half timg [128][64]                  9216
constexpr int V2_H2 = V2_H1 + 9216;                     // H2 bimg [64][256] -> dH2 bimg -> X^T timg [160][64]            18432
constexpr int V2_STAGE = V2_H2 + 18432;
constexpr int V2_DOUT = V2_STAGE + NS2 * CHUNK_F;       // dOut image [64][16]
constexpr int V2_OUTH = V2_DOUT + 1024;
constexpr int V2_DPL = V2_OUTH;                         // dOut plain [64][12] aliases the head outputs (dead once q_a / dOut are formed)
constexpr int V2_SMALL = V2_OUTH + 1024;
constexpr int V2_INT = V2_SMALL + SM_SMALL_N;
constexpr int V2_FLOATS = V2_INT + 2 * 256;
constexpr size_t TC2_SMEM = sizeof(float) * V2_FLOATS + 8 * (2 * NS2 + 3) + 16;
static_assert(TC2_SMEM <= 227 * 1024, "shared memory budget");

// batch-major image [64 rows][K cols] with the padded chunk stride: consecutive features of one batch row sit in
// consecutive banks, so a warp whose lanes own consecutive features scatters one batch column conflict-free
__device__ __forceinline__ int bimg_off(int b, int c, int K) { return (b >> 3) * ((K >> 2) * TP_CH) + (c >> 2) * TP_CH + (b & 7) * 4 + (c & 3); }
__device__ __forceinline__ uint64_t desc_bimg(uint32_t saddr, int K) { return make_desc(saddr, TP_CH * 4, (uint32_t)(K >> 2) * TP_CH * 4); }

// the MMAs of one weight chunk: MH feature halves x KS k-steps, fully unrolled (straight-line, descriptors by constant adds)
template <int MH, int KS>
__device__ __forceinline__ void chunk_mmas(uint32_t d_tmem, int n, uint64_t a_desc, uint32_t a_half_inc, uint64_t b_desc, uint32_t b_inc,
                                           uint32_t idesc, uint32_t acc) {
#pragma unroll
    for (int h = 0; h < MH; ++h)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
            mma_tf32(d_tmem + h * n, a_desc + h * a_half_inc + ks * 16u, b_desc + ks * b_inc, idesc, acc | (uint32_t)(ks != 0));
}

constexpr int NTHREADS2 = NEPI + 96;     // warps 0-7 epilogue, warp 8 weight-stream producer, warps 9-10 MMA issuers

// chunk schedule of one event in the issuer's consumption order (the eval L1^T is issued right after the target L2^T)
__device__ __forceinline__ void sched_entry2(int i, int& net, int& chunk) {
    if (i < 5) { net = 0; chunk = WI_W1 + i; }                  // target L1^T
    else if (i < 13) { net = 0; chunk = WI_W2K + (i - 5); }     // target L2^T
    else if (i < 18) { net = 1; chunk = WI_W1 + (i - 13); }     // eval L1^T
    else if (i == 18) { net = 0; chunk = WI_WH; }               // target head
    else if (i < 27) { net = 1; chunk = WI_W2K + (i - 19); }    // eval L2^T
    else if (i == 27) { net = 1; chunk = WI_WH; }               // eval head
    else if (i == 28) { net = 1; chunk = WI_WHT; }              // dH2^T
    else { net = 1; chunk = WI_W2T + (i - 29); }                // dH1^T
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Roles: the producer (warp 8) streams the fixed 37-chunk schedule of every event through the ring; the issuer (warp 9)
// runs the fixed stage list of an event, each stage released by a `go` arrival of all 256 epilogue threads (their
// shared-memory images are written and fenced) and reported back through `done` (main chain) or `doneL1` (the L1^T
// stages, which run ahead: the eval L1^T under the target L2 epilogue, the NEXT event's target L1^T under the dW1
// epilogue).  The epilogue warps therefore never wait for weight chunks, only for finished accumulators.
__global__ void __launch_bounds__(NTHREADS2, 1) k_learn_dueling_tc2(const TcLearnParams P) {
    using L = Layout<RL_MODEL_DUELING>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw);
    float* sX = sm + V2_X; float* sH1 = sm + V2_H1; float* sH2 = sm + V2_H2;
    float* sH1T = sX;          // H1^T / dH1^T (feature-major) live in the X region once the eval L1 MMAs are done
    float* sDT = sH1;          // dH2^T half buffer lives in the H1 region once the eval L2 MMAs are done
    float* sXT = sH2;          // X^T lives in the dH2 region once the dH1 MMAs are done
    float* sStage = sm + V2_STAGE; float* sDout = sm + V2_DOUT; float* sOuth = sm + V2_OUTH; float* sDpl = sm + V2_DPL;
    float* nq = sm + V2_SMALL + 128; float* gb = nq + 64; float* red = gb + 64;
    float* bias_t = red + 32;                 // b1[128] b2[256] bh[16] of the target net
    float* bias_e = bias_t + 400;             // same for the eval net
    int* meta = reinterpret_cast<int*>(sm + V2_INT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + V2_FLOATS);
    uint64_t* full = bars; uint64_t* empty = bars + NS2; uint64_t* done = bars + 2 * NS2; uint64_t* doneL1 = done + 1; uint64_t* go = done + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* Pe = P.lb.params; const float* Pt = P.lb.target;
    float* G = P.lb.grad_scratch + (size_t)blockIdx.x * L::N_TRAIN;
    const int total = *P.ev_total;
    const int n_my = total > (int)blockIdx.x ? (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }     // a slot is released by both issuers
        mbar_init(done, 2); mbar_init(doneL1, 1); mbar_init(go, NEPI);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < NEPI) {
        for (int i = threadIdx.x; i < 400; i += NEPI) {
            const int o = i < 128 ? L::OFF_B1 + i : i < 384 ? L::OFF_B2 + (i - 128) : L::OFF_BH + (i - 384);
            const bool ok = i < 384 + 9;
            bias_t[i] = ok ? Pt[o] : 0.f; bias_e[i] = ok ? Pe[o] : 0.f;
        }
        for (int i = threadIdx.x; i < L::N_TRAIN / 4; i += NEPI) reinterpret_cast<float4*>(G)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t T_WORK = tmem, T_DW2 = tmem + 256;
    const uint32_t T_L1 = tmem + 192;          // L1^T accumulator [128][64] has its own columns (dW1^T uses 0-159)
    const int S = P.cfg.slot_cap, cap = P.rp.capacity;
    const uint32_t aX = smem_u32(sX), aH1 = smem_u32(sH1), aH1T = smem_u32(sH1T), aH2 = smem_u32(sH2), aD = smem_u32(sDout), aXT = smem_u32(sXT);

    if (warp == 8) {
        // =================================== weight-stream producer ===================================
        if (lane == 0) {
            const uint32_t n_chunks = (uint32_t)n_my * SCHED_N;
            for (uint32_t produced = 0; produced < n_chunks; ++produced) {
                const uint32_t slot = produced % NS2;
                if (produced >= NS2) mbar_wait(&empty[slot], ((produced / NS2) - 1) & 1);
                int net, ch; sched_entry2(produced % SCHED_N, net, ch);
                bulk_load(sStage + slot * CHUNK_F, (net ? P.wimg_e : P.wimg_t) + (size_t)ch * CHUNK_F, CHUNK_F * 4, &full[slot]);
            }
        }
    } else if (warp == 9 || warp == 10) {
        // =================================== MMA issuers (two warps) ===================================
        // A tcgen05.mma costs its issuing thread ~100 cycles whatever its N (rl_tc_mma_bench), and two threads issue
        // concurrently at ~1.5x the rate, so the long stages are split between two issuers: the 128-feature halves of
        // L2^T / dH2^T, and the K range of the head and of dH1^T (second half into its own accumulator columns, summed by
        // the epilogue).  Both issuers walk the same go sequence and the same chunk stream; a chunk slot is released by
        // two arrivals (tcgen05.commit of each issuer that read it, plain arrive of the one that did not), `done` likewise.
        if (lane == 0) {
            const int role = warp - 9;
            uint32_t consumed = 0, go_no = 0;
            auto wait_go = [&]() { mbar_wait(go, go_no & 1); ++go_no; fence_after(); };
            auto chunk_wait = [&]() -> uint32_t {
                const uint32_t slot = consumed % NS2;
                mbar_wait(&full[slot], (consumed / NS2) & 1);
                fence_after();
                return smem_u32(sStage + slot * CHUNK_F);
            };
            auto chunk_release = [&]() { mma_commit(&empty[consumed % NS2]); ++consumed; };
            // a chunk this issuer does not read: wait until it has landed (so that this arrival cannot be counted for the slot's
            // PREVIOUS round, which the other issuer may still be reading), then release this issuer's share of the slot
            auto chunk_skip = [&](int n) { for (int i = 0; i < n; ++i) { (void)chunk_wait(); mbar_arrive(&empty[consumed % NS2]); ++consumed; } };
            auto mma_seq = [&](uint32_t d_tmem, uint64_t a_desc, uint32_t a_inc, uint64_t b_desc, uint32_t b_inc, uint32_t idesc, int n, uint32_t acc) {
#pragma unroll 1
                for (int i = 0; i < n; ++i) {
                    mma_tf32(d_tmem, a_desc, b_desc, idesc, acc);
                    a_desc += a_inc; b_desc += b_inc; acc = 1u;
                }
            };
            // L1^T (issuer 0): 5 chunks [128][32], 4 k-steps each, B = X image
            auto l1 = [&]() {
                if (role == 1) { chunk_skip(5); return; }
                const uint32_t idesc = make_idesc(128, 64, 0, 0);
                uint64_t b_desc = make_desc(aX, 128u, RL_K1 * 32u);
#pragma unroll 1
                for (int c = 0; c < 5; ++c) {
                    chunk_mmas<1, 4>(T_L1, 64, desc_kmajor(chunk_wait(), 32), 0u, b_desc, 16u, idesc, c != 0);
                    b_desc += 64u;
                    chunk_release();
                }
                mma_commit(doneL1);
            };
            // L2^T: 8 chunks [256][16]; issuer r takes feature half r (accumulator columns 64r..), B = H1 bimg
            auto l2 = [&]() {
                const uint32_t idesc = make_idesc(128, 64, 0, 0);
                uint64_t b_desc = make_desc(aH1, TP_CH * 4u, 32u * TP_CH * 4u);
                const uint32_t a_off = role ? (128u * 16u * 4u) : 0u;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    chunk_mmas<1, 2>(T_WORK + role * 64, 64, desc_kmajor(chunk_wait() + a_off, 16), 0u, b_desc, TP_KSTEP >> 4, idesc, c != 0);
                    b_desc += 2u * (TP_KSTEP >> 4);
                    chunk_release();
                }
                mma_commit(done);
            };
            // head (M = 64): issuer r takes k-steps 16r..16r+15 into accumulator columns 16r.. (the epilogue adds the two)
            auto head = [&]() {
                const uint32_t b_base = chunk_wait();
                mma_seq(T_WORK + role * 16, desc_bimg(aH2 + role * 16 * TP_KSTEP, 256), TP_KSTEP >> 4, desc_kmajor(b_base + role * 16 * 256, 256), 16u,
                        make_idesc(64, 16, 0, 0), 16, 0u);
                chunk_release();
                mma_commit(done);
            };
            for (int it = 0; it < n_my; ++it) {
                if (it == 0) { wait_go(); l1(); }                 // target L1^T of event 0 (later ones are issued one event ahead)
                wait_go(); l2(); l1();                            // target L2^T, then the eval L1^T (runs under the target L2 epilogue)
                wait_go(); head();                                // target head
                wait_go(); l2();                                  // eval L2^T
                wait_go(); head();                                // eval head
                wait_go();                                        // dH2^T[n2][b] = Wh[n2][:] . dOut[b][:]: issuer r takes feature half r
                {
                    const uint64_t a_desc = desc_kmajor(chunk_wait() + (role ? 128u * 16u * 4u : 0u), 16);
                    mma_seq(T_WORK + role * 64, a_desc, 16u, desc_kmajor(aD, 16), 16u, make_idesc(128, 64, 0, 0), 2, 0u);
                    chunk_release();
                    mma_commit(done);
                }
                wait_go();                                        // dW2 half 0 (TMEM-resident accumulator): issuer 0
                if (role == 0) {
                    mma_seq(T_DW2, desc_timg(aH1T), TP_KSTEP >> 4, desc_timg(aH1), TP_KSTEP >> 4, make_idesc(128, 128, 0, 0), 8, it != 0);
                    mma_commit(done);
                } else {
                    mbar_arrive(done);
                }
                wait_go();                                        // dW2 half 1 (issuer 0); dH1^T[k1][b] = W2^T[k1][:] . dH2[b][:] split over K:
                {                                                 // of every [128][32] chunk issuer r takes k-steps 2r, 2r+1 into columns 64r..
                    if (role == 0)
                        mma_seq(T_DW2 + 128, desc_timg(aH1T), TP_KSTEP >> 4, desc_timg(aH1), TP_KSTEP >> 4, make_idesc(128, 128, 0, 0), 8, it != 0);
                    const uint32_t idesc = make_idesc(128, 64, 0, 0);
                    uint64_t b_desc = make_desc(aH2 + role * 2 * TP_KSTEP, TP_CH * 4u, 64u * TP_CH * 4u);
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        chunk_mmas<1, 2>(T_WORK + role * 64, 64, desc_kmajor(chunk_wait() + role * 2 * 256, 32), 0u, b_desc, TP_KSTEP >> 4, idesc, c != 0);
                        b_desc += 4u * (TP_KSTEP >> 4);
                        chunk_release();
                    }
                    mma_commit(done);
                }
                wait_go();                                        // dW1^T = dH1^T X: issuer 0
                if (role == 0) {
                    mma_seq(T_WORK, desc_timg(aH1T), TP_KSTEP >> 4, desc_timg(aXT), TP_KSTEP >> 4, make_idesc(128, 160, 0, 0), 8, 0u);
                    mma_commit(done);
                } else {
                    mbar_arrive(done);
                }
                if (it + 1 < n_my) { wait_go(); l1(); }           // next event's target L1^T (runs under the dW1 epilogue)
            }
        }
    } else {
        // =================================== epilogue warps ===================================
        uint32_t done_no = 0, l1_no = 0;
        const int q = warp & 3, half = warp >> 2;
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const int f1 = q * 32 + lane;                   // feature owned in 128-feature stages (k1)
        const int f2 = half * 128 + f1;                 // feature owned in 256-feature stages (n2) == threadIdx.x
        const int row64 = q * 16 + lane;                // M = 64 accumulator (head only): rows 16q+i in lanes 32q+i, i < 16
        const bool rvalid = lane < 16;
        // this thread's shared-memory images are written: make them visible to the tensor core and release the next stage
        auto go_signal = [&]() { fence_proxy_async(); fence_before(); mbar_arrive(go); };
        auto wait_done = [&]() { mbar_wait(done, done_no & 1); ++done_no; fence_after(); };
        auto wait_l1 = [&]() { mbar_wait(doneL1, l1_no & 1); ++l1_no; fence_after(); };
        int tr_n = 0;
        auto stamp = [&](int it) { if (P.trace && blockIdx.x == 0 && threadIdx.x == 0 && it < 8 && tr_n < 30) P.trace[it * 40 + tr_n++] = clock64(); };
        int meta_i = 0; size_t meta_ring = 0;
        auto load_meta_a = [&](int b, int e) {           // phase A (warps 6-7): sampled ring positions
            if (threadIdx.x >= NEPI - R) {
                const int r = threadIdx.x - (NEPI - R);
                meta_ring = (size_t)(P.ev_rows[e] / S) * cap;
                meta_i = max(P.sample_idx[(size_t)e * R + r], 0);     // -1 = skipped by the uniform sampler (error state)
                meta[b * 256 + r] = meta_i;
            }
        };
        auto load_meta_b = [&](int b) {                  // phase B: action / reward / done of those positions
            if (threadIdx.x >= NEPI - R) {
                const int r = threadIdx.x - (NEPI - R);
                int* m = meta + b * 256;
                m[64 + r] = P.rp.action[meta_ring + meta_i];
                reinterpret_cast<float*>(m)[128 + r] = P.rp.reward[meta_ring + meta_i];
                reinterpret_cast<float*>(m)[192 + r] = (float)P.rp.done[meta_ring + meta_i];
            }
        };
        auto prefetch_rows = [&](size_t rg_, const int* ids) {
            for (int v = threadIdx.x; v < 2 * R * 5; v += NEPI) {
                const int which = v / (R * 5), rem = v - which * (R * 5), r = rem / 5, ln = rem - r * 5;
                const float* p = (which ? P.rp.obs : P.rp.next_obs) + (rg_ + ids[r]) * RL_K1 + ln * 32;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
        };
        float4 xr[10];
        // EVENT-list rows of the current / next / next-next event: loaded early, divided (ring = row / S * cap) only where used,
        // so that no thread stalls on the load
        int row_cur = 0, row_n1 = 0, row_n2 = 0;
        if (n_my > 0) {
            // prologue: metadata + target-net input of event 0, its target L1^T released, eval-net rows of event 0 and the
            // ring positions of event 1 on their way
            load_meta_a(0, blockIdx.x);
            load_meta_b(0);
            epi_bar();
            row_cur = P.ev_rows[blockIdx.x];
            const size_t ring0 = (size_t)(row_cur / S) * cap;
            gather_load(xr, P.rp.next_obs + ring0 * RL_K1, meta);
            gather_store<false>(sX, xr);
            go_signal();                                                                     // -> target L1^T
            gather_load(xr, P.rp.obs + ring0 * RL_K1, meta);
            if (n_my > 1) { load_meta_a(1, blockIdx.x + gridDim.x); row_n1 = P.ev_rows[blockIdx.x + gridDim.x]; }
            epi_bar();          // event 1's ring positions (written by warps 6-7) are read by every warp's prefetch_rows below
        }
        // L1 epilogue: lane = feature k1, this warp's 32 batch columns; H1 = relu(D + b1)
        auto l1_epilogue = [&](const float* bias, bool eval) {
            float v[32];
            tmem_ld32(T_L1 + t_lane + half * 32, v);
            tmem_wait_ld();
            const float b1 = bias[f1];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = to_tf32(fmaxf(v[j] + b1, 0.f));
#pragma unroll
            for (int j = 0; j < 32; ++j) sH1[bimg_off(half * 32 + j, f1, 128)] = v[j];                // batch-major: B of L2^T
            if (eval) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)                                                         // feature-major: A of dW2, relu mask of dH1
                    *reinterpret_cast<float4*>(sH1T + timg_off(f1, half * 32 + j4 * 4)) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            }
        };
        // L2 epilogue: lane = feature n2, all 64 batch columns; H2 = relu(D + b2) -> batch-major image
        auto l2_epilogue = [&](const float* bias) {
            const float b2 = bias[128 + f2];
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
                float v[32];
                tmem_ld32(T_WORK + t_lane + half * 64 + cb * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) sH2[bimg_off(cb * 32 + j, f2, 256)] = to_tf32(fmaxf(v[j] + b2, 0.f));
            }
        };
        // head epilogue: [A(8) | V] + bh -> sOuth; returns the mean of the whole [64, 8] advantage tensor
        auto head_epilogue = [&](const float* bias) -> float {
            if (half == 0) {
                float v[16], v2[16];
                tmem_ld16(T_WORK + t_lane, v);                 // k-steps 0-15 (issuer 0)
                tmem_ld16(T_WORK + t_lane + 16, v2);           // k-steps 16-31 (issuer 1)
                tmem_wait_ld();
                if (rvalid) {
#pragma unroll
                    for (int j = 0; j < 9; ++j) sOuth[row64 * 16 + j] = (v[j] + v2[j]) + bias[384 + j];
                }
            }
            fence_before();
            epi_bar();
            float s = 0.f;
            for (int o = threadIdx.x; o < R * 8; o += NEPI) s += sOuth[(o >> 3) * 16 + (o & 7)];
            return epi_sum(s, red) * (1.0f / (8 * R));
        };
        for (int it = 0; it < n_my; ++it) {
            tr_n = 0; stamp(it);
            const int e = blockIdx.x + it * gridDim.x;
            const bool more = it + 1 < n_my, more2 = it + 2 < n_my;
            const int* idx = meta + (it & 1) * 256; const int* act = idx + 64;
            const float* rew = reinterpret_cast<const float*>(idx + 128); const float* dn = rew + 64;
            const size_t ring = (size_t)(row_cur / S) * cap;
            const size_t ring_n1 = (size_t)(row_n1 / S) * cap;
            // ---------------- target net ----------------
            wait_l1(); stamp(it);                                   // target L1^T (issued one event ahead)
            gather_store<false>(sX, xr);                            // eval-net input; sX is free: the target L1 MMAs have completed
            l1_epilogue(bias_t, false);
            go_signal();                                            // -> target L2^T, eval L1^T (two `go` arrivals may never follow each
                                                                    //    other without a `done` wait in between: parity waits would alias)
            if (more) {                                             // hidden behind them
                load_meta_b((it + 1) & 1);
                prefetch_rows(ring_n1, meta + ((it + 1) & 1) * 256);
            }
            wait_done(); stamp(it);
            l2_epilogue(bias_t);
            go_signal();                                            // -> target head
            wait_done(); stamp(it);
            {
                const float mean = head_epilogue(bias_t);
                if (threadIdx.x < R) {
                    const float* o = sOuth + threadIdx.x * 16;
                    float mx = o[0];
#pragma unroll
                    for (int j = 1; j < 8; ++j) mx = fmaxf(mx, o[j]);
                    nq[threadIdx.x] = mx + o[8] - mean;
                }
                epi_bar();
            }
            // ---------------- eval net ----------------
            wait_l1(); stamp(it);
            l1_epilogue(bias_e, true);
            go_signal();                                            // -> eval L2^T
            wait_done(); stamp(it);
            l2_epilogue(bias_e);
            go_signal();                                            // -> eval head
            wait_done(); stamp(it);
            const float mean_e = head_epilogue(bias_e);
            stamp(it);
            // ---- TD target, loss, priorities, dOut ----
            {
                float g = 0.f, sq = 0.f;
                if (threadIdx.x < R) {
                    const int b = threadIdx.x;
                    const float qa = sOuth[b * 16 + act[b]] + sOuth[b * 16 + 8] - mean_e;
                    const float y = rew[b] + P.lb.gamma * (1.0f - dn[b]) * nq[b];
                    const float diff = qa - y;
                    g = 2.0f * diff * (1.0f / R);
                    sq = diff * diff;
                    gb[b] = g;
                    P.lb.new_prio[(size_t)e * R + b] = fabsf(nq[b] - qa);
                }
                const float gsum = epi_sum(g, red);
                const float loss = epi_sum(sq, red) * (1.0f / R);
                if (threadIdx.x == 0) P.lb.loss[e] = loss;
                const float shift = gsum * (1.0f / (8 * R));
                for (int o = threadIdx.x; o < R * 16; o += NEPI) {
                    const int b = o >> 4, j = o & 15;
                    const float d = j < 8 ? ((j == act[b] ? gb[b] : 0.f) - shift) : (j == 8 ? gb[b] : 0.f);
                    sDout[img_off(b, j, 16)] = to_tf32(d);
                    if (j < 12) sDpl[b * 12 + j] = d;
                }
            }
            go_signal();                                            // -> dH2^T (runs under the head-gradient SIMT below)
            epi_bar();                                              // sDpl complete for every warp
            stamp(it);
            // ---- head gradients (SIMT): dWh[k][j] += sum_b H2[b][k] dOut[b][j]; dbh ----
            {
                const int k = threadIdx.x;
                float acc[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) acc[j] = 0.f;
                const float* hk = sH2 + (k >> 2) * TP_CH + (k & 3);
#pragma unroll 2
                for (int g = 0; g < 8; ++g) {
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int b = g * 8 + rr;
                        const float h = hk[g * (64 * TP_CH) + rr * 4];
                        const float4 d0 = *reinterpret_cast<const float4*>(sDpl + b * 12), d1 = *reinterpret_cast<const float4*>(sDpl + b * 12 + 4);
                        const float d8 = sDpl[b * 12 + 8];
                        acc[0] = fmaf(h, d0.x, acc[0]); acc[1] = fmaf(h, d0.y, acc[1]); acc[2] = fmaf(h, d0.z, acc[2]); acc[3] = fmaf(h, d0.w, acc[3]);
                        acc[4] = fmaf(h, d1.x, acc[4]); acc[5] = fmaf(h, d1.y, acc[5]); acc[6] = fmaf(h, d1.z, acc[6]); acc[7] = fmaf(h, d1.w, acc[7]);
                        acc[8] = fmaf(h, d8, acc[8]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 9; ++j) red_add(G + L::OFF_WH + k * 9 + j, acc[j]);
                if (threadIdx.x < 9) {
                    float s = 0.f;
                    for (int b = 0; b < R; ++b) s += sDpl[b * 12 + threadIdx.x];
                    red_add(G + L::OFF_BH + threadIdx.x, s);
                }
            }
            // ---- dH2 epilogue: lane = feature n2 (the column this thread just read in the SIMT loop); mask by H2 > 0;
            //      dH2 batch-major in place (B of dH1^T), dH2^T feature-major (B of dW2) through the half buffer:
            //      features 0-127 now, 128-255 after the first dW2 half has been consumed ----
            wait_done(); stamp(it);
            float dv[64];
            {
                float sb2 = 0.f;
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
                    float v[32];
                    tmem_ld32(T_WORK + t_lane + half * 64 + cb * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float* ph = sH2 + bimg_off(cb * 32 + j, f2, 256);
                        const float m = *ph > 0.f ? to_tf32(v[j]) : 0.f;
                        *ph = m;
                        dv[cb * 32 + j] = m;
                        sb2 += m;
                    }
                }
                red_add(G + L::OFF_B2 + f2, sb2);                                  // db2[n2] = sum_b dH2[b][n2]
                if (half == 0) {
#pragma unroll
                    for (int j4 = 0; j4 < 16; ++j4)
                        *reinterpret_cast<float4*>(sDT + timg_off(f1, j4 * 4)) = make_float4(dv[j4 * 4], dv[j4 * 4 + 1], dv[j4 * 4 + 2], dv[j4 * 4 + 3]);
                }
            }
            go_signal();                                            // -> dW2 half 0
            wait_done(); stamp(it);                                 // it finished reading the half buffer
            if (half == 1) {
#pragma unroll
                for (int j4 = 0; j4 < 16; ++j4)
                    *reinterpret_cast<float4*>(sDT + timg_off(f1, j4 * 4)) = make_float4(dv[j4 * 4], dv[j4 * 4 + 1], dv[j4 * 4 + 2], dv[j4 * 4 + 3]);
            }
            go_signal();                                            // -> dW2 half 1, dH1^T
            gather_load(xr, P.rp.obs + ring * RL_K1, idx);          // X rows again (for the X^T image), hidden behind the dH1 MMAs
            wait_done(); stamp(it);
            {   // dH1 epilogue: lane = feature k1, this warp's 32 batch columns; mask by H1 > 0, dH1^T in place of H1^T, db1
                float v[32], v2[32];
                tmem_ld32(T_WORK + t_lane + half * 32, v);          // k-steps 0,1 of every chunk (issuer 0)
                tmem_ld32(T_WORK + t_lane + 64 + half * 32, v2);    // k-steps 2,3 (issuer 1)
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += v2[j];
                float sb1 = 0.f;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float4* ph = reinterpret_cast<float4*>(sH1T + timg_off(f1, half * 32 + j4 * 4));
                    const float4 h = *ph;
                    float4 o;
                    o.x = h.x > 0.f ? to_tf32(v[j4 * 4 + 0]) : 0.f; o.y = h.y > 0.f ? to_tf32(v[j4 * 4 + 1]) : 0.f;
                    o.z = h.z > 0.f ? to_tf32(v[j4 * 4 + 2]) : 0.f; o.w = h.w > 0.f ? to_tf32(v[j4 * 4 + 3]) : 0.f;
                    *ph = o;
                    sb1 += (o.x + o.y) + (o.z + o.w);
                }
                red_add(G + L::OFF_B1 + f1, sb1);                                  // db1[k1] (two warps share a feature)
            }
            stamp(it);
            gather_store<true>(sXT, xr);                            // X^T image for dW1^T (dH2 region is dead: dH1 MMAs are done)
            stamp(it);
            go_signal();                                            // -> dW1^T
            stamp(it);
            if (more) gather_load(xr, P.rp.next_obs + ring_n1 * RL_K1, meta + ((it + 1) & 1) * 256);
            stamp(it);
            wait_done(); stamp(it);
            if (more) {
                gather_store<false>(sX, xr);                        // sX (dH1^T) is free: the dW1 MMAs have completed
                stamp(it);
                go_signal();                                        // -> next event's target L1^T (runs under the dW1 epilogue below)
                stamp(it);
                gather_load(xr, P.rp.obs + ring_n1 * RL_K1, meta + ((it + 1) & 1) * 256);   // its eval-net rows
                stamp(it);
                if (more2) {                                        // ring positions of the event after it (this event's buffer is dead)
                    load_meta_a(it & 1, e + 2 * gridDim.x);
                    row_n2 = P.ev_rows[e + 2 * gridDim.x];
                }
                stamp(it);
            }
            {   // dW1^T flush: the accumulator has one k1 row per lane, so a direct flush touches 32 different 128-byte lines per
                // instruction.  Transpose through the (now dead) X^T / dH2 region instead: rows are staged with an odd leading
                // dimension (conflict-free column writes), then read back row-wise so that a warp adds 128 contiguous bytes.
                float* stg = sH2;
                float* gw = G + L::OFF_W1T;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const int c_lo = pass ? 96 : 0, ncol = pass ? 64 : 96, ldw = ncol + 1;
                    for (int cb = (pass ? 3 : 0) + half; cb < (pass ? 5 : 3); cb += 2) {
                        float v[32];
                        tmem_ld32(T_WORK + t_lane + cb * 32, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) stg[f1 * ldw + (cb * 32 - c_lo) + j] = v[j];
                    }
                    epi_bar();
                    for (int i = threadIdx.x; i < 128 * ncol; i += NEPI) {
                        const int r = pass ? (i >> 6) : (int)__umulhi((uint32_t)i, 44739243u), c = i - r * ncol;      // i / 96 by multiply-high
                        red_add(gw + r * RL_K1 + c_lo + c, stg[r * ldw + c]);
                    }
                    if (pass == 0) epi_bar();
                }
            }
            stamp(it);
            fence_before();
            epi_bar();
            row_cur = row_n1; row_n1 = row_n2;
            stamp(it);
        }
        if (n_my > 0) {     // flush the TMEM-resident dW2 accumulator once (the dW1 `done` covered every earlier MMA)
            for (int cb = 0; cb < 4; ++cb) {
                const int c0 = half * 128 + cb * 32;
                float v[32];
                tmem_ld32(T_DW2 + t_lane + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    *reinterpret_cast<float4*>(G + L::OFF_W2T + f1 * 256 + c0 + j4 * 4) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

// =====================================================================================================
// k_learn_dueling_h -- the event kernel of k_learn_dueling_tc2 on fp16 operands (tcgen05.mma kind::f16, fp32
// accumulation in TMEM).  fp16 carries the same 11 significant bits as tf32, but an instruction covers K = 16 instead of
// 8 and every operand byte count halves: 114 MMA instructions per event instead of 228 (the tensor pipe is
// instruction-issue bound at N = 64, see rl_tc_mma_bench), 296 KB of weights per event instead of 592 KB, and the
// weight chunks of a whole stage fit in the ring.  Geometry in BYTES is the one of the tf32 kernel (16-byte chunks = 8
// halves, 128-byte core matrices, padded chunk stride 144 B), so roles, stages, barriers and TMEM map are unchanged.
// The backward operands (dOut, dH2, dH1) are scaled by H_SCALE = 256 (exact power of two, removed when the fp32
// gradients leave TMEM) so that small gradients stay out of the fp16 subnormal range; values are clamped to +-60000.
// =====================================================================================================
constexpr int NSH = 8;                                  // weight-chunk stages of 8 KB
constexpr int HCHUNK = 4096;                            // halves per weight chunk (8 KB): same element shapes as the tf32 chunks
constexpr int HC = 72;                                  // halves between consecutive 16-byte chunks in padded images (144 B)
constexpr float H_SCALE = 256.0f;
// byte offsets of the regions
constexpr int HB_X = 0;                                 // X' / X plain [64][160] (20480 B) -> H1^T / dH1^T timg [128][64] (18432 B)
constexpr int HB_H1 = HB_X + 20480;                     // H1 bimg [64][128] (18432 B) -> full dH2^T timg [256][64] (36864 B)
constexpr int HB_H2 = HB_H1 + 36864;                    // H2 bimg [64][256] (36864 B) -> dH2 bimg -> X^T timg [160][64] (23040 B) -> fp32 dW1 staging (<= 49664 B)
constexpr int HB_XT = HB_H2 + 49920;                    // X^T timg [160][64] (23040 B), written once per event from the gathered eval rows
constexpr int HB_STAGE = HB_XT + 23040;
constexpr int HB_DOUT = HB_STAGE + NSH * HCHUNK * 2;    // dOut image [64][16] halves (2048 B)
constexpr int HB_DOUTT = HB_DOUT + 2048;                // dOut^T image [16][64] halves, padded chunk stride (2304 B): B of the dWh GEMM
constexpr int HB_OUTH = HB_DOUTT + 2304;                // head outputs [64][16] floats (4096 B); dOut plain [64][12] aliases it
constexpr int HB_SMALL = HB_OUTH + 4096;                // nq, gb, red, biases (floats)
constexpr int HB_INT = HB_SMALL + 4 * SM_SMALL_N;
constexpr int HB_BARS = HB_INT + 4 * 512;
constexpr size_t TCH_SMEM = HB_BARS + 8 * (2 * NSH + 3) + 16;
static_assert(TCH_SMEM <= 227 * 1024 && HB_BARS % 8 == 0, "shared memory budget");

__device__ __forceinline__ int himg_off(int r, int c, int K) { return (r >> 3) * (K * 8) + (c >> 3) * 64 + (r & 7) * 8 + (c & 7); }
__device__ __forceinline__ int hbimg_off(int b, int c, int K) { return (b >> 3) * ((K >> 3) * HC) + (c >> 3) * HC + (b & 7) * 8 + (c & 7); }
__device__ __forceinline__ int htimg_off(int f, int b) { return (f >> 3) * (8 * HC) + (b >> 3) * HC + (f & 7) * 8 + (b & 7); }
__device__ __forceinline__ uint64_t desc_hplain(uint32_t saddr, int K) { return make_desc(saddr, 128u, (uint32_t)K * 16u); }
__device__ __forceinline__ uint64_t desc_hbimg(uint32_t saddr, int K) { return make_desc(saddr, 144u, (uint32_t)(K >> 3) * 144u); }
__device__ __forceinline__ uint64_t desc_htimg(uint32_t saddr) { return make_desc(saddr, 144u, 8u * 144u); }
__host__ __device__ constexpr uint32_t make_idesc_h(int M, int N) {      // kind::f16: A, B fp16 (format 0), D fp32
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ __half h_sat(float x) { return __float2half_rn(fminf(fmaxf(x, -60000.f), 60000.f)); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {      // one F2FP.PACK_AB (full rate) instead of two F2F (quarter rate, XU pipe)
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));     // upper half <- b, lower half <- a
    return r;
}
__device__ __forceinline__ uint32_t pack_h2_sat(float a, float b) { return (uint32_t)__half_as_ushort(h_sat(a)) | ((uint32_t)__half_as_ushort(h_sat(b)) << 16); }
__device__ __forceinline__ float h_lo(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xFFFFu))); }
__device__ __forceinline__ float h_hi(uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); }

// store the gathered 64 x 160 fp32 rows as an fp16 image (plain [64][160] or feature-major X^T [160][64])
template <bool TRANSPOSED>
__device__ __forceinline__ void gather_store_h(__half* img, const float4 (&x)[10]) {
#pragma unroll
    for (int u = 0; u < 10; ++u) {
        const int v = threadIdx.x + u * NEPI;
        const int rr = v & 7, cc = (v >> 3) & 3, blk = v >> 5;
        const int rg = blk / 10, cg = blk - rg * 10;
        const int r = rg * 8 + rr, c = (cg * 4 + cc) * 4;
        if (!TRANSPOSED) {
            *reinterpret_cast<uint2*>(img + himg_off(r, c, RL_K1)) = make_uint2(pack_h2(x[u].x, x[u].y), pack_h2(x[u].z, x[u].w));
        } else {
            img[htimg_off(c + 0, r)] = __float2half_rn(x[u].x); img[htimg_off(c + 1, r)] = __float2half_rn(x[u].y);
            img[htimg_off(c + 2, r)] = __float2half_rn(x[u].z); img[htimg_off(c + 3, r)] = __float2half_rn(x[u].w);
        }
    }
}

// the MMAs of one fp16 weight chunk: KS k-steps of 16, straight-line
template <int KS>
__device__ __forceinline__ void chunk_mmas_h(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t b_inc, uint32_t idesc, uint32_t acc) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) mma_f16(d_tmem, a_desc + ks * 16u, b_desc + ks * b_inc, idesc, acc | (uint32_t)(ks != 0));
}

struct TcLearnParamsH {
    TcLearnParams p;
    const __half* wimg_e;
    const __half* wimg_t;
};

__global__ void __launch_bounds__(NTHREADS2, 1) k_learn_dueling_h(const TcLearnParamsH PH) {
    using L = Layout<RL_MODEL_DUELING>;
    const TcLearnParams& P = PH.p;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __half* sX = reinterpret_cast<__half*>(smem_raw + HB_X);
    __half* sH1 = reinterpret_cast<__half*>(smem_raw + HB_H1);
    __half* sH2 = reinterpret_cast<__half*>(smem_raw + HB_H2);
    __half* sH1T = sX;         // H1^T / dH1^T (feature-major) live in the X region once the eval L1 MMAs are done
    // (H2^T, then the full dH2^T image [256][64], live in the double-size H1 region once the eval L2 MMAs are done)
    __half* sXT = reinterpret_cast<__half*>(smem_raw + HB_XT);
    __half* sStage = reinterpret_cast<__half*>(smem_raw + HB_STAGE);
    __half* sDout = reinterpret_cast<__half*>(smem_raw + HB_DOUT);
    __half* sDoutT = reinterpret_cast<__half*>(smem_raw + HB_DOUTT);
    float* sOuth = reinterpret_cast<float*>(smem_raw + HB_OUTH); float* sDpl = sOuth;
    float* smallf = reinterpret_cast<float*>(smem_raw + HB_SMALL);
    float* nq = smallf + 128; float* gb = nq + 64; float* red = gb + 64;
    float* bias_t = red + 32; float* bias_e = bias_t + 400;
    int* meta = reinterpret_cast<int*>(smem_raw + HB_INT);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + HB_BARS);
    uint64_t* full = bars; uint64_t* empty = bars + NSH; uint64_t* done = bars + 2 * NSH; uint64_t* doneL1 = done + 1; uint64_t* go = done + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* Pe = P.lb.params; const float* Pt = P.lb.target;
    float* G = P.lb.grad_scratch + (size_t)blockIdx.x * L::N_TRAIN;
    const int total = *P.ev_total;
    const int n_my = total > (int)blockIdx.x ? (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSH; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }
        mbar_init(done, 2); mbar_init(doneL1, 1); mbar_init(go, NEPI);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x < NEPI) {
        for (int i = threadIdx.x; i < 400; i += NEPI) {
            const int o = i < 128 ? L::OFF_B1 + i : i < 384 ? L::OFF_B2 + (i - 128) : L::OFF_BH + (i - 384);
            const bool ok = i < 384 + 9;
            bias_t[i] = ok ? Pt[o] : 0.f; bias_e[i] = ok ? Pe[o] : 0.f;
        }
        for (int i = threadIdx.x; i < L::N_TRAIN / 4; i += NEPI) reinterpret_cast<float4*>(G)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t T_WORK = tmem, T_DW2 = tmem + 256, T_L1 = tmem + 192;
    const int S = P.cfg.slot_cap, cap = P.rp.capacity;
    const uint32_t aX = smem_u32(sX), aH1 = smem_u32(sH1), aH1T = smem_u32(sH1T), aH2 = smem_u32(sH2), aD = smem_u32(sDout), aXT = smem_u32(sXT), aDT2 = smem_u32(sDoutT);

    if (warp == 8) {
        if (lane == 0) {
            const uint32_t n_chunks = (uint32_t)n_my * SCHED_N;
            for (uint32_t produced = 0; produced < n_chunks; ++produced) {
                const uint32_t slot = produced % NSH;
                if (produced >= NSH) mbar_wait(&empty[slot], ((produced / NSH) - 1) & 1);
                int net, ch; sched_entry2(produced % SCHED_N, net, ch);
                bulk_load(sStage + slot * HCHUNK, (net ? PH.wimg_e : PH.wimg_t) + (size_t)ch * HCHUNK, HCHUNK * 2, &full[slot]);
            }
        }
    } else if (warp == 9 || warp == 10) {
        if (lane == 0) {
            const int role = warp - 9;
            uint32_t consumed = 0, go_no = 0;
            auto wait_go = [&]() { mbar_wait(go, go_no & 1); ++go_no; fence_after(); };
            auto chunk_wait = [&]() -> uint32_t {
                const uint32_t slot = consumed % NSH;
                mbar_wait(&full[slot], (consumed / NSH) & 1);
                fence_after();
                return smem_u32(sStage + slot * HCHUNK);
            };
            auto chunk_release = [&]() { mma_commit(&empty[consumed % NSH]); ++consumed; };
            auto chunk_skip = [&](int n) { for (int i = 0; i < n; ++i) { (void)chunk_wait(); mbar_arrive(&empty[consumed % NSH]); ++consumed; } };
            auto mma_seq = [&](uint32_t d_tmem, uint64_t a_desc, uint32_t a_inc, uint64_t b_desc, uint32_t b_inc, uint32_t idesc, int n, uint32_t acc) {
#pragma unroll 1
                for (int i = 0; i < n; ++i) {
                    mma_f16(d_tmem, a_desc, b_desc, idesc, acc);
                    a_desc += a_inc; b_desc += b_inc; acc = 1u;
                }
            };
            // L1^T (issuer 0): 5 chunks [128][32]: 2 k-steps of 16 each, B = X plain image
            auto l1 = [&]() {
                if (role == 1) { chunk_skip(5); return; }
                const uint32_t idesc = make_idesc_h(128, 64);
                uint64_t b_desc = desc_hplain(aX, RL_K1);
#pragma unroll 1
                for (int c = 0; c < 5; ++c) {
                    chunk_mmas_h<2>(T_L1, desc_hplain(chunk_wait(), 32), b_desc, 16u, idesc, c != 0);
                    b_desc += 32u;
                    chunk_release();
                }
                mma_commit(doneL1);
            };
            // L2^T: 8 chunks [256][16]: one k-step; issuer r takes feature half r, B = H1 bimg
            auto l2 = [&]() {
                const uint32_t idesc = make_idesc_h(128, 64);
                uint64_t b_desc = desc_hbimg(aH1, 128);
                const uint32_t a_off = role ? 4096u : 0u;               // rows 128..255 of a [256][16] fp16 chunk
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {
                    chunk_mmas_h<1>(T_WORK + role * 64, desc_hplain(chunk_wait() + a_off, 16), b_desc, 18u, idesc, c != 0);
                    b_desc += 18u;
                    chunk_release();
                }
                mma_commit(done);
            };
            // head (M = 64): 16 k-steps, issuer r takes 8r..8r+7 into accumulator columns 16r..
            auto head = [&]() {
                const uint32_t b_base = chunk_wait();
                mma_seq(T_WORK + role * 16, desc_hbimg(aH2 + role * 8 * 288, 256), 18u, desc_hplain(b_base + role * 8 * 256, 256), 16u,
                        make_idesc_h(64, 16), 8, 0u);
                chunk_release();
                mma_commit(done);
            };
            for (int it = 0; it < n_my; ++it) {
                if (it == 0) { wait_go(); l1(); }
                wait_go(); l2(); l1();
                wait_go(); head();
                wait_go(); l2();
                wait_go(); head();
                wait_go();                                        // dH2^T: one k-step, issuer r takes feature half r
                {
                    mma_f16(T_WORK + role * 64, desc_hplain(chunk_wait() + (role ? 4096u : 0u), 16), desc_hplain(aD, 16), make_idesc_h(128, 64), 0u);
                    chunk_release();
                    // dWh[n2][j] = sum_b H2[b][n2] dOut[b][j]: A = H2^T rows of feature half r (in the H1 region), B = dOut^T, K = 64 batch rows
                    mma_seq(T_WORK + 128 + role * 16, desc_htimg(aH1 + role * 16 * 1152), 18u, desc_htimg(aDT2), 18u, make_idesc_h(128, 16), 4, 0u);
                    mma_commit(done);
                }
                wait_go();                                        // dW2 = H1^T dH2 (issuer 0: M = 128, N = 256, K = 64 batch rows = 4 k-steps, TMEM-resident
                {                                                 // accumulator); dH1^T split over K: of every [128][32] chunk issuer r takes k-step r
                    if (role == 0) mma_seq(T_DW2, desc_htimg(aH1T), 18u, desc_htimg(aH1), 18u, make_idesc_h(128, 256), 4, it != 0);
                    const uint32_t idesc = make_idesc_h(128, 64);
                    uint64_t b_desc = desc_hbimg(aH2 + role * 288, 256);
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        chunk_mmas_h<1>(T_WORK + role * 64, desc_hplain(chunk_wait() + role * 256, 32), b_desc, 18u, idesc, c != 0);
                        b_desc += 36u;
                        chunk_release();
                    }
                    mma_commit(done);
                }
                wait_go();                                        // dW1^T = dH1^T X: 4 k-steps (issuer 0)
                if (role == 0) {
                    mma_seq(T_WORK, desc_htimg(aH1T), 18u, desc_htimg(aXT), 18u, make_idesc_h(128, 160), 4, 0u);
                    mma_commit(done);
                } else {
                    mbar_arrive(done);
                }
                if (it + 1 < n_my) { wait_go(); l1(); }
            }
        }
    } else {
        // =================================== epilogue warps ===================================
        uint32_t done_no = 0, l1_no = 0;
        const int q = warp & 3, half = warp >> 2;
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const int f1 = q * 32 + lane;
        const int f2 = half * 128 + f1;
        const int row64 = q * 16 + lane;
        const bool rvalid = lane < 16;
        auto go_signal = [&]() { fence_proxy_async(); fence_before(); mbar_arrive(go); };
        auto wait_done = [&]() { mbar_wait(done, done_no & 1); ++done_no; fence_after(); };
        auto wait_l1 = [&]() { mbar_wait(doneL1, l1_no & 1); ++l1_no; fence_after(); };
        int meta_i = 0; size_t meta_ring = 0;
        auto load_meta_a = [&](int b, int e) {
            if (threadIdx.x >= NEPI - R) {
                const int r = threadIdx.x - (NEPI - R);
                meta_ring = (size_t)(P.ev_rows[e] / S) * cap;
                meta_i = max(P.sample_idx[(size_t)e * R + r], 0);
                meta[b * 256 + r] = meta_i;
            }
        };
        auto load_meta_b = [&](int b) {
            if (threadIdx.x >= NEPI - R) {
                const int r = threadIdx.x - (NEPI - R);
                int* m = meta + b * 256;
                m[64 + r] = P.rp.action[meta_ring + meta_i];
                reinterpret_cast<float*>(m)[128 + r] = P.rp.reward[meta_ring + meta_i];
                reinterpret_cast<float*>(m)[192 + r] = (float)P.rp.done[meta_ring + meta_i];
            }
        };
        auto prefetch_rows = [&](size_t rg_, const int* ids) {
            for (int v = threadIdx.x; v < 2 * R * 5; v += NEPI) {
                const int which = v / (R * 5), rem = v - which * (R * 5), r = rem / 5, ln = rem - r * 5;
                const float* p = (which ? P.rp.obs : P.rp.next_obs) + (rg_ + ids[r]) * RL_K1 + ln * 32;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
        };
        float4 xr[10];
        int row_cur = 0, row_n1 = 0, row_n2 = 0;
        if (n_my > 0) {
            load_meta_a(0, blockIdx.x);
            load_meta_b(0);
            epi_bar();
            row_cur = P.ev_rows[blockIdx.x];
            const size_t ring0 = (size_t)(row_cur / S) * cap;
            gather_load(xr, P.rp.next_obs + ring0 * RL_K1, meta);
            gather_store_h<false>(sX, xr);
            go_signal();
            gather_load(xr, P.rp.obs + ring0 * RL_K1, meta);
            if (n_my > 1) { load_meta_a(1, blockIdx.x + gridDim.x); row_n1 = P.ev_rows[blockIdx.x + gridDim.x]; }
            epi_bar();          // event 1's ring positions (written by warps 6-7) are read by every warp's prefetch_rows below
        }
        // L1 epilogue: lane = feature k1, this warp's 32 batch columns; H1 = relu(D + b1)
        auto l1_epilogue = [&](const float* bias, bool eval) {
            float v[32];
            tmem_ld32(T_L1 + t_lane + half * 32, v);
            tmem_wait_ld();
            const float b1 = bias[f1];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + b1, 0.f);
#pragma unroll
            for (int j = 0; j < 32; ++j) sH1[hbimg_off(half * 32 + j, f1, 128)] = __float2half_rn(v[j]);      // batch-major: B of L2^T
            if (eval) {
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8)                                                                   // feature-major: A of dW2, relu mask of dH1
                    *reinterpret_cast<uint4*>(sH1T + htimg_off(f1, half * 32 + j8 * 8)) =
                        make_uint4(pack_h2(v[j8 * 8], v[j8 * 8 + 1]), pack_h2(v[j8 * 8 + 2], v[j8 * 8 + 3]), pack_h2(v[j8 * 8 + 4], v[j8 * 8 + 5]), pack_h2(v[j8 * 8 + 6], v[j8 * 8 + 7]));
            }
        };
        auto l2_epilogue = [&](const float* bias, bool eval) {
            const float b2 = bias[128 + f2];
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
                float v[32];
                tmem_ld32(T_WORK + t_lane + half * 64 + cb * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + b2, 0.f);
#pragma unroll
                for (int j = 0; j < 32; ++j) sH2[hbimg_off(cb * 32 + j, f2, 256)] = __float2half_rn(v[j]);      // batch-major: A of the head GEMM
                if (eval) {                        // feature-major H2^T [256][64] in the (dead) H1 region: A of the dWh GEMM, relu mask of dH2
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8)
                        *reinterpret_cast<uint4*>(sH1 + htimg_off(f2, cb * 32 + j8 * 8)) =
                            make_uint4(pack_h2(v[j8 * 8], v[j8 * 8 + 1]), pack_h2(v[j8 * 8 + 2], v[j8 * 8 + 3]), pack_h2(v[j8 * 8 + 4], v[j8 * 8 + 5]), pack_h2(v[j8 * 8 + 6], v[j8 * 8 + 7]));
                }
            }
        };
        auto head_epilogue = [&](const float* bias) -> float {
            if (half == 0) {
                float v[16], v2[16];
                tmem_ld16(T_WORK + t_lane, v);
                tmem_ld16(T_WORK + t_lane + 16, v2);
                tmem_wait_ld();
                if (rvalid) {
#pragma unroll
                    for (int j = 0; j < 9; ++j) sOuth[row64 * 16 + j] = (v[j] + v2[j]) + bias[384 + j];
                }
            }
            fence_before();
            epi_bar();
            float s = 0.f;
            for (int o = threadIdx.x; o < R * 8; o += NEPI) s += sOuth[(o >> 3) * 16 + (o & 7)];
            return epi_sum(s, red) * (1.0f / (8 * R));
        };
        for (int it = 0; it < n_my; ++it) {
            const int e = blockIdx.x + it * gridDim.x;
            const bool more = it + 1 < n_my, more2 = it + 2 < n_my;
            const int* idx = meta + (it & 1) * 256; const int* act = idx + 64;
            const float* rew = reinterpret_cast<const float*>(idx + 128); const float* dn = rew + 64;
            const size_t ring_n1 = (size_t)(row_n1 / S) * cap;
            // ---------------- target net ----------------
            wait_l1();
            gather_store_h<false>(sX, xr);
            l1_epilogue(bias_t, false);
            go_signal();                                            // -> target L2^T, eval L1^T
            gather_store_h<true>(sXT, xr);                          // X^T image for this event's dW1^T, from the same registers (hidden behind L2^T)
            if (more) {
                load_meta_b((it + 1) & 1);
                prefetch_rows(ring_n1, meta + ((it + 1) & 1) * 256);
            }
            wait_done();
            l2_epilogue(bias_t, false);
            go_signal();                                            // -> target head
            wait_done();
            {
                const float mean = head_epilogue(bias_t);
                if (threadIdx.x < R) {
                    const float* o = sOuth + threadIdx.x * 16;
                    float mx = o[0];
#pragma unroll
                    for (int j = 1; j < 8; ++j) mx = fmaxf(mx, o[j]);
                    nq[threadIdx.x] = mx + o[8] - mean;
                }
                epi_bar();
            }
            // ---------------- eval net ----------------
            wait_l1();
            l1_epilogue(bias_e, true);
            go_signal();                                            // -> eval L2^T
            wait_done();
            l2_epilogue(bias_e, true);
            go_signal();                                            // -> eval head
            wait_done();
            const float mean_e = head_epilogue(bias_e);
            // ---- TD target, loss, priorities, dOut (fp16 image scaled by H_SCALE, fp32 plain copy unscaled) ----
            {
                float g = 0.f, sq = 0.f;
                if (threadIdx.x < R) {
                    const int b = threadIdx.x;
                    const float qa = sOuth[b * 16 + act[b]] + sOuth[b * 16 + 8] - mean_e;
                    const float y = rew[b] + P.lb.gamma * (1.0f - dn[b]) * nq[b];
                    const float diff = qa - y;
                    g = 2.0f * diff * (1.0f / R);
                    sq = diff * diff;
                    gb[b] = g;
                    P.lb.new_prio[(size_t)e * R + b] = fabsf(nq[b] - qa);
                }
                const float gsum = epi_sum(g, red);
                const float loss = epi_sum(sq, red) * (1.0f / R);
                if (threadIdx.x == 0) P.lb.loss[e] = loss;
                const float shift = gsum * (1.0f / (8 * R));
                for (int o = threadIdx.x; o < R * 16; o += NEPI) {
                    const int b = o >> 4, j = o & 15;
                    const float d = j < 8 ? ((j == act[b] ? gb[b] : 0.f) - shift) : (j == 8 ? gb[b] : 0.f);
                    const __half dh = h_sat(d * H_SCALE);
                    sDout[himg_off(b, j, 16)] = dh;
                    sDoutT[htimg_off(j, b)] = dh;
                    if (j < 12) sDpl[b * 12 + j] = d;
                }
            }
            go_signal();                                            // -> dH2^T and dWh (tensor core)
            epi_bar();
            if (threadIdx.x < 9) {                                  // dbh[j] = sum_b dOut[b][j]
                float s = 0.f;
                for (int b = 0; b < R; ++b) s += sDpl[b * 12 + threadIdx.x];
                red_add(G + L::OFF_BH + threadIdx.x, s);
            }
            wait_done();
            {   // head weight gradients from the tensor core: row n2 = f2 of dWh (scaled by H_SCALE)
                float wv[16];
                tmem_ld16(T_WORK + t_lane + 128 + half * 16, wv);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 9; ++j) red_add(G + L::OFF_WH + f2 * 9 + j, wv[j] * (1.0f / H_SCALE));
            }
            // ---- dH2 epilogue (values carry the factor H_SCALE): lane = feature n2; mask by H2 > 0 read from H2^T; dH2 batch-major
            //      over H2 (B of dH1^T), dH2^T feature-major in place of H2^T (B of dW2) ----
            {
                float sb2 = 0.f;
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
                    float v[32];
                    tmem_ld32(T_WORK + t_lane + half * 64 + cb * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        uint4* pt = reinterpret_cast<uint4*>(sH1 + htimg_off(f2, cb * 32 + j8 * 8));
                        const uint4 hq = *pt;
                        const uint32_t hw[4] = {hq.x, hq.y, hq.z, hq.w};
                        uint32_t ow[4];
#pragma unroll
                        for (int p2 = 0; p2 < 4; ++p2) {
                            const int j = j8 * 8 + p2 * 2;
                            const float m0 = h_lo(hw[p2]) > 0.f ? v[j] : 0.f, m1 = h_hi(hw[p2]) > 0.f ? v[j + 1] : 0.f;
                            const __half q0 = h_sat(m0), q1 = h_sat(m1);
                            sH2[hbimg_off(cb * 32 + j, f2, 256)] = q0;
                            sH2[hbimg_off(cb * 32 + j + 1, f2, 256)] = q1;
                            ow[p2] = (uint32_t)__half_as_ushort(q0) | ((uint32_t)__half_as_ushort(q1) << 16);
                            sb2 += m0 + m1;
                        }
                        *pt = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    }
                }
                red_add(G + L::OFF_B2 + f2, sb2 * (1.0f / H_SCALE));
            }
            go_signal();                                            // -> dW2, dH1^T
            if (more) gather_load(xr, P.rp.next_obs + ring_n1 * RL_K1, meta + ((it + 1) & 1) * 256);   // next event's target-net rows, two stages ahead
            wait_done();
            {   // dH1 epilogue: lane = feature k1, this warp's 32 batch columns; mask by H1 > 0, dH1^T in place of H1^T, db1
                float v[32], v2[32];
                tmem_ld32(T_WORK + t_lane + half * 32, v);
                tmem_ld32(T_WORK + t_lane + 64 + half * 32, v2);
                tmem_wait_ld();
                float sb1 = 0.f;
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    uint4* ph = reinterpret_cast<uint4*>(sH1T + htimg_off(f1, half * 32 + j8 * 8));
                    const uint4 h = *ph;
                    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
                    uint32_t ow[4];
#pragma unroll
                    for (int p2 = 0; p2 < 4; ++p2) {
                        const int j = j8 * 8 + p2 * 2;
                        const float a = h_lo(hw[p2]) > 0.f ? v[j] + v2[j] : 0.f, b = h_hi(hw[p2]) > 0.f ? v[j + 1] + v2[j + 1] : 0.f;
                        sb1 += a + b;
                        ow[p2] = pack_h2_sat(a, b);
                    }
                    *ph = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
                red_add(G + L::OFF_B1 + f1, sb1 * (1.0f / H_SCALE));
            }
            go_signal();                                            // -> dW1^T
            wait_done();
            if (more) {
                gather_store_h<false>(sX, xr);
                go_signal();                                        // -> next event's target L1^T
                gather_load(xr, P.rp.obs + ring_n1 * RL_K1, meta + ((it + 1) & 1) * 256);
                if (more2) {
                    load_meta_a(it & 1, e + 2 * gridDim.x);
                    row_n2 = P.ev_rows[e + 2 * gridDim.x];
                }
            }
            {   // dW1^T flush through a shared-memory transpose (see k_learn_dueling_tc2), unscaled on the way
                float* stg = reinterpret_cast<float*>(sH2);
                float* gw = G + L::OFF_W1T;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const int c_lo = pass ? 96 : 0, ncol = pass ? 64 : 96, ldw = ncol + 1;
                    for (int cb = (pass ? 3 : 0) + half; cb < (pass ? 5 : 3); cb += 2) {
                        float v[32];
                        tmem_ld32(T_WORK + t_lane + cb * 32, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) stg[f1 * ldw + (cb * 32 - c_lo) + j] = v[j] * (1.0f / H_SCALE);
                    }
                    epi_bar();
                    for (int i = threadIdx.x; i < 128 * ncol; i += NEPI) {
                        const int r = pass ? (i >> 6) : (int)__umulhi((uint32_t)i, 44739243u), c = i - r * ncol;
                        red_add(gw + r * RL_K1 + c_lo + c, stg[r * ldw + c]);
                    }
                    if (pass == 0) epi_bar();
                }
            }
            fence_before();
            epi_bar();
            row_cur = row_n1; row_n1 = row_n2;
        }
        if (n_my > 0) {     // flush the TMEM-resident dW2 accumulator once, unscaled
            for (int cb = 0; cb < 4; ++cb) {
                const int c0 = half * 128 + cb * 32;
                float v[32];
                tmem_ld32(T_DW2 + t_lane + c0, v);
                tmem_wait_ld();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4)
                    *reinterpret_cast<float4*>(G + L::OFF_W2T + f1 * 256 + c0 + j4 * 4) =
                        make_float4(v[j4 * 4] * (1.0f / H_SCALE), v[j4 * 4 + 1] * (1.0f / H_SCALE), v[j4 * 4 + 2] * (1.0f / H_SCALE), v[j4 * 4 + 3] * (1.0f / H_SCALE));
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 512);
}

// =====================================================================================================
// k_act_dueling_h -- brain.get_action for the dueling brains with the structure of k_learn_dueling_h: fp16 operands,
// transposed-output GEMMs (M = 128 features, N = 64 rows of a tile of the brain's ALL list), producer warp + two MMA
// issuers + eight epilogue warps, the NEXT tile's L1^T issued one tile ahead (its rows are gathered behind the L2^T stage).
// Epilogue = k_brain_act's: per-row dueling combine at B = 1 (PERD3QN.py:202), first-max argmax, exploration draws keyed
// (t_act, slot) (PERD3QN.py:204-210), rec[].action.
// =====================================================================================================
struct TcActParams {
    rl_world_cfg cfg;
    rl_agent_rec* rec;
    const float* obs;          // obs_state
    const int32_t* rows;       // row list of this brain, kind ALL
    const int32_t* total;      // device scalar
    const float* params;       // biases are read from the kernel-layout buffer
    const float* wimg;
    const double* epsilon;
    uint64_t t_act;
    float* q_out;              // [row_cap][8] or null
};
constexpr int ACT_SCHED_N = 14;
constexpr int AB_X = 0;                                 // X plain [64][160] (20480 B)
constexpr int AB_H1 = AB_X + 20480;                     // H1 bimg [64][128] (18432 B)
constexpr int AB_H2 = AB_H1 + 18432;                    // H2 bimg [64][256] (36864 B)
constexpr int AB_STAGE = AB_H2 + 36864;
constexpr int AB_BIAS = AB_STAGE + NSH * HCHUNK * 2;    // b1[128] b2[256] bh[16] floats
constexpr int AB_IDS = AB_BIAS + 4 * 400;               // [2][64] row ids
constexpr int AB_BARS = AB_IDS + 4 * 128;
constexpr size_t ACTH_SMEM = AB_BARS + 8 * (2 * NSH + 3) + 16;
static_assert(ACTH_SMEM <= 227 * 1024 && AB_BARS % 8 == 0, "shared memory budget");

struct TcActParamsH {
    TcActParams p;
    const __half* wimg;
};

__global__ void __launch_bounds__(NTHREADS2, 1) k_act_dueling_h(const TcActParamsH PH) {
    using L = Layout<RL_MODEL_DUELING>;
    const TcActParams& P = PH.p;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __half* sX = reinterpret_cast<__half*>(smem_raw + AB_X);
    __half* sH1 = reinterpret_cast<__half*>(smem_raw + AB_H1);
    __half* sH2 = reinterpret_cast<__half*>(smem_raw + AB_H2);
    __half* sStage = reinterpret_cast<__half*>(smem_raw + AB_STAGE);
    float* bias = reinterpret_cast<float*>(smem_raw + AB_BIAS);
    int* ids = reinterpret_cast<int*>(smem_raw + AB_IDS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + AB_BARS);
    uint64_t* full = bars; uint64_t* empty = bars + NSH; uint64_t* done = bars + 2 * NSH; uint64_t* doneL1 = done + 1; uint64_t* go = done + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = *P.total;
    const int n_tiles = (total + R - 1) / R;
    const int n_my = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSH; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }
        mbar_init(done, 2); mbar_init(doneL1, 1); mbar_init(go, NEPI);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 256);
    if (threadIdx.x < NEPI)
        for (int i = threadIdx.x; i < 400; i += NEPI) {
            const int o = i < 128 ? L::OFF_B1 + i : i < 384 ? L::OFF_B2 + (i - 128) : L::OFF_BH + (i - 384);
            bias[i] = i < 384 + 9 ? P.params[o] : 0.f;
        }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t T_WORK = tmem, T_L1 = tmem + 192;
    const int S = P.cfg.slot_cap;
    const uint32_t aX = smem_u32(sX), aH1 = smem_u32(sH1), aH2 = smem_u32(sH2);

    if (warp == 8) {
        if (lane == 0) {
            const uint32_t n_chunks = (uint32_t)n_my * ACT_SCHED_N;
            for (uint32_t produced = 0; produced < n_chunks; ++produced) {
                const uint32_t slot = produced % NSH;
                if (produced >= NSH) mbar_wait(&empty[slot], ((produced / NSH) - 1) & 1);
                // consumption order of a tile: L2^T (W2K x 8), head (WH), then the NEXT tile's L1^T (W1 x 5); the very first
                // tile's L1^T chunks lead the stream
                const uint32_t i = produced % ACT_SCHED_N;
                const int ch = i < 5 ? WI_W1 + (int)i : i < 13 ? WI_W2K + (int)(i - 5) : WI_WH;
                bulk_load(sStage + slot * HCHUNK, PH.wimg + (size_t)ch * HCHUNK, HCHUNK * 2, &full[slot]);
            }
        }
    } else if (warp == 9 || warp == 10) {
        if (lane == 0) {
            const int role = warp - 9;
            uint32_t consumed = 0, go_no = 0;
            auto wait_go = [&]() { mbar_wait(go, go_no & 1); ++go_no; fence_after(); };
            auto chunk_wait = [&]() -> uint32_t {
                const uint32_t slot = consumed % NSH;
                mbar_wait(&full[slot], (consumed / NSH) & 1);
                fence_after();
                return smem_u32(sStage + slot * HCHUNK);
            };
            auto chunk_release = [&]() { mma_commit(&empty[consumed % NSH]); ++consumed; };
            auto chunk_skip = [&](int n) { for (int i = 0; i < n; ++i) { (void)chunk_wait(); mbar_arrive(&empty[consumed % NSH]); ++consumed; } };
            auto l1 = [&]() {
                if (role == 1) { chunk_skip(5); return; }
                const uint32_t idesc = make_idesc_h(128, 64);
                uint64_t b_desc = desc_hplain(aX, RL_K1);
#pragma unroll 1
                for (int c = 0; c < 5; ++c) {
                    chunk_mmas_h<2>(T_L1, desc_hplain(chunk_wait(), 32), b_desc, 16u, idesc, c != 0);
                    b_desc += 32u;
                    chunk_release();
                }
                mma_commit(doneL1);
            };
            for (int it = 0; it < n_my; ++it) {
                if (it == 0) { wait_go(); l1(); }
                wait_go();                                        // L2^T: issuer r takes feature half r
                {
                    const uint32_t idesc = make_idesc_h(128, 64);
                    uint64_t b_desc = desc_hbimg(aH1, 128);
                    const uint32_t a_off = role ? 4096u : 0u;
#pragma unroll 1
                    for (int c = 0; c < 8; ++c) {
                        chunk_mmas_h<1>(T_WORK + role * 64, desc_hplain(chunk_wait() + a_off, 16), b_desc, 18u, idesc, c != 0);
                        b_desc += 18u;
                        chunk_release();
                    }
                    mma_commit(done);
                }
                wait_go();                                        // head (M = 64): issuer r takes k-steps 8r..8r+7 into columns 16r..
                {
                    const uint32_t b_base = chunk_wait();
                    const uint32_t idesc = make_idesc_h(64, 16);
                    uint64_t a_desc = desc_hbimg(aH2 + role * 8 * 288, 256), b_desc = desc_hplain(b_base + role * 8 * 256, 256);
#pragma unroll 1
                    for (int ks = 0; ks < 8; ++ks) { mma_f16(T_WORK + role * 16, a_desc, b_desc, idesc, ks != 0); a_desc += 18u; b_desc += 16u; }
                    chunk_release();
                    mma_commit(done);
                }
                if (it + 1 < n_my) { wait_go(); l1(); }           // next tile's L1^T (runs under this tile's head epilogue)
            }
        }
    } else {
        uint32_t done_no = 0, l1_no = 0;
        const int q = warp & 3, half = warp >> 2;
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const int f1 = q * 32 + lane, f2 = half * 128 + f1;
        const int row64 = q * 16 + lane;
        const bool rvalid = lane < 16;
        auto go_signal = [&]() { fence_proxy_async(); fence_before(); mbar_arrive(go); };
        auto wait_done = [&]() { mbar_wait(done, done_no & 1); ++done_no; fence_after(); };
        auto wait_l1 = [&]() { mbar_wait(doneL1, l1_no & 1); ++l1_no; fence_after(); };
        auto load_ids = [&](int b, int tile) {           // rows past `total` read row 0 (results discarded)
            if (threadIdx.x >= NEPI - R) {
                const int r = threadIdx.x - (NEPI - R), i = tile * R + r;
                ids[b * R + r] = i < total ? P.rows[i] : 0;
            }
        };
        float4 xr[10];
        if (n_my > 0) {
            load_ids(0, blockIdx.x);
            if (n_my > 1) load_ids(1, blockIdx.x + gridDim.x);
            epi_bar();
            gather_load(xr, P.obs, ids);
            gather_store_h<false>(sX, xr);
            go_signal();                                          // -> L1^T of the first tile
        }
        const double epsilon = *P.epsilon;
        for (int it = 0; it < n_my; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const bool more = it + 1 < n_my;
            // this tile's row id is read before the first `go` arrival: the id buffer is recycled (for tile it+2) only by threads
            // that have seen the L2^T `done`, i.e. after all 256 arrivals
            const int rid = ids[(it & 1) * R + (row64 & 63)];
            wait_l1();
            {   // L1 epilogue: lane = feature k1, this warp's 32 rows; H1 = relu(D + b1) -> batch-major image
                float v[32];
                tmem_ld32(T_L1 + t_lane + half * 32, v);
                tmem_wait_ld();
                const float b1 = bias[f1];
#pragma unroll
                for (int j = 0; j < 32; ++j) sH1[hbimg_off(half * 32 + j, f1, 128)] = __float2half_rn(fmaxf(v[j] + b1, 0.f));
            }
            go_signal();                                          // -> L2^T
            if (more) gather_load(xr, P.obs, ids + ((it + 1) & 1) * R);   // next tile's rows, hidden behind L2^T and the head
            wait_done();
            {   // L2 epilogue: lane = feature n2, all 64 rows; H2 = relu(D + b2) -> batch-major image
                const float b2 = bias[128 + f2];
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
                    float v[32];
                    tmem_ld32(T_WORK + t_lane + half * 64 + cb * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) sH2[hbimg_off(cb * 32 + j, f2, 256)] = __float2half_rn(fmaxf(v[j] + b2, 0.f));
                }
            }
            go_signal();                                          // -> head
            if (more) {                                           // sX is free since this tile's L1^T completed
                gather_store_h<false>(sX, xr);
                if (it + 2 < n_my) load_ids(it & 1, tile + 2 * gridDim.x);
            }
            wait_done();
            if (more) go_signal();                                // -> next tile's L1^T (runs under the head epilogue below)
            if (half == 0) {
                float v[16], v2[16];
                tmem_ld16(T_WORK + t_lane, v);
                tmem_ld16(T_WORK + t_lane + 16, v2);
                tmem_wait_ld();
                const int i = tile * R + row64;
                if (rvalid && i < total) {
                    float qv[8], ssum = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { qv[j] = (v[j] + v2[j]) + bias[384 + j]; ssum += qv[j]; }
                    const float val = (v[8] + v2[8]) + bias[384 + 8], mean = ssum * 0.125f;     // B = 1: per-row mean (PERD3QN.py:202)
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
            fence_before();
            epi_bar();
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem, 256);
}

// fp16 weight images: the chunk shapes and order of the tf32 images, 4096 halves (8 KB) per chunk
__global__ void k_build_wimg_dueling_h(const float* __restrict__ p, __half* __restrict__ wimg) {
    using L = Layout<RL_MODEL_DUELING>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 128 * 160) {
        const int n = i / 160, k = i - n * 160;
        wimg[(WI_W1 + k / 32) * HCHUNK + himg_off(n, k % 32, 32)] = __float2half_rn(p[L::OFF_W1T + k * 128 + n]);
    }
    if (i < 128 * 256) {
        const int k1 = i / 256, n2 = i - k1 * 256;
        const __half v = __float2half_rn(p[L::OFF_W2T + i]);
        wimg[(WI_W2K + k1 / 16) * HCHUNK + himg_off(n2, k1 % 16, 16)] = v;
        wimg[(WI_W2T + n2 / 32) * HCHUNK + himg_off(k1, n2 % 32, 32)] = v;
        wimg[(WI_W2N + (n2 / 128) * 4 + k1 / 32) * HCHUNK + himg_off(n2 % 128, k1 % 32, 32)] = v;
    }
    if (i < 256 * 16) {
        const int n2 = i / 16, j = i - n2 * 16;
        const __half v = j < 9 ? __float2half_rn(p[L::OFF_WH + n2 * 9 + j]) : __float2half_rn(0.f);
        wimg[WI_WH * HCHUNK + himg_off(j, n2, 256)] = v;
        wimg[WI_WHT * HCHUNK + himg_off(n2, j, 16)] = v;
    }
}

// =====================================================================================================
// brain.get_action for the dueling brains on the tensor cores: the forward half of the event kernel over 64-row
// tiles of the brain's ALL row list (Helpers/trainer.py:88-89; PERD3QN.py:81-89,198-210; D3QN.py:82-93,161-173).
// Same pipeline: warp 8 streams the 14 weight chunks of a tile (W1[5], W2K[8], WH), epilogue thread 0 issues the
// MMAs, the next tile's observation rows are gathered into registers behind the L2 stage.  The epilogue is the one of
// k_brain_act (per-row dueling combine at B = 1, first-max argmax, exploration draws keyed (t_act, slot)).
// =====================================================================================================

__global__ void __launch_bounds__(NTHREADS, 1) k_act_dueling_tc(const TcActParams P) {
    using L = Layout<RL_MODEL_DUELING>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sm = reinterpret_cast<float*>(smem_raw);
    float* sX = sm + SM_X; float* sH1 = sm + SM_H1; float* sH2 = sm + SM_H2;
    float* sStage = sm + SM_STAGE;
    float* bias = sm + SM_SMALL;                           // b1[128] b2[256] bh[16]
    int* ids = reinterpret_cast<int*>(sm + SM_INT);        // [2][64] row ids of the current / next tile
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_FLOATS);
    uint64_t* full = bars; uint64_t* empty = bars + NS; uint64_t* done = bars + 2 * NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = *P.total;
    const int n_tiles = (total + R - 1) / R;
    const int n_my = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 256);
    if (threadIdx.x < NEPI)
        for (int i = threadIdx.x; i < 400; i += NEPI) {
            const int o = i < 128 ? L::OFF_B1 + i : i < 384 ? L::OFF_B2 + (i - 128) : L::OFF_BH + (i - 384);
            bias[i] = i < 384 + 9 ? P.params[o] : 0.f;
        }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t T_WORK = *tmem_slot;
    const int S = P.cfg.slot_cap;

    if (warp == 8) {
        if (lane == 0) {
            const uint32_t n_chunks = (uint32_t)n_my * ACT_SCHED_N;
            for (uint32_t produced = 0; produced < n_chunks; ++produced) {
                const uint32_t slot = produced % NS;
                if (produced >= NS) mbar_wait(&empty[slot], ((produced / NS) - 1) & 1);
                bulk_load(sStage + slot * CHUNK_F, P.wimg + (size_t)(produced % ACT_SCHED_N) * CHUNK_F, CHUNK_F * 4, &full[slot]);
            }
        }
    } else {
        uint32_t stage_no = 0, consumed = 0;
        const int q = warp & 3, half = warp >> 2;
        const uint32_t t_lane = (uint32_t)(q * 32) << 16;
        const int row = q * 16 + lane;
        const bool rvalid = lane < 16;
        const uint32_t aX = smem_u32(sX), aH1 = smem_u32(sH1), aH2 = smem_u32(sH2);
        auto stage_sync = [&]() { fence_proxy_async(); fence_before(); epi_bar(); };
        auto stream_gemm = [&](uint32_t d_tmem, uint32_t a_base, int a_k, int nch, int kc, int m, int n) {
            const uint32_t idesc = make_idesc(m, n, 0, 0);
            for (int c = 0; c < nch; ++c) {
                const uint32_t slot = consumed % NS;
                mbar_wait(&full[slot], (consumed / NS) & 1);
                fence_after();
                const uint32_t b_base = smem_u32(sStage + slot * CHUNK_F);
                for (int ks = 0; ks < kc / 8; ++ks) {
                    const int kcol = c * kc + ks * 8;
                    mma_tf32(d_tmem, desc_kmajor(a_base + (kcol >> 2) * 128, a_k), desc_kmajor(b_base + ks * 256, kc), idesc, (c | ks) != 0);
                }
                mma_commit(&empty[slot]);
                ++consumed;
            }
        };
        auto wait_done = [&]() { mbar_wait(done, stage_no & 1); ++stage_no; fence_after(); };
        auto load_ids = [&](int b, int tile) {           // rows past `total` read row 0 (results discarded)
            if (threadIdx.x >= NEPI - R) {
                const int r = threadIdx.x - (NEPI - R), i = tile * R + r;
                ids[b * R + r] = i < total ? P.rows[i] : 0;
            }
        };
        float4 xr[10];
        if (n_my > 0) {
            load_ids(0, blockIdx.x);
            epi_bar();
            gather_load(xr, P.obs, ids);
            gather_store<false>(sX, xr);
        }
        const double epsilon = *P.epsilon;
        for (int it = 0; it < n_my; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const bool more = it + 1 < n_my;
            const int* idc = ids + (it & 1) * R;
            RL_STAGE(stream_gemm(T_WORK, aX, RL_K1, 5, 32, 64, 128));
            if (more) load_ids((it + 1) & 1, tile + gridDim.x);          // hidden behind the L1 MMAs
            wait_done();
            {   // L1 epilogue: H1 = relu(D + b1)
                float va[2][32];
                tmem_ld32(T_WORK + t_lane + half * 64, va[0]);
                tmem_ld32(T_WORK + t_lane + half * 64 + 32, va[1]);
                tmem_wait_ld();
                if (rvalid) {
#pragma unroll
                    for (int cb = 0; cb < 2; ++cb) {
                        const int c0 = half * 64 + cb * 32;
                        const float* v = va[cb];
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                            *reinterpret_cast<float4*>(sH1 + img_off(row, c0 + j4 * 4, 128)) =
                                make_float4(to_tf32(fmaxf(v[j4 * 4] + bias[c0 + j4 * 4], 0.f)), to_tf32(fmaxf(v[j4 * 4 + 1] + bias[c0 + j4 * 4 + 1], 0.f)),
                                            to_tf32(fmaxf(v[j4 * 4 + 2] + bias[c0 + j4 * 4 + 2], 0.f)), to_tf32(fmaxf(v[j4 * 4 + 3] + bias[c0 + j4 * 4 + 3], 0.f)));
                    }
                }
            }
            RL_STAGE(stream_gemm(T_WORK, aH1, 128, 8, 16, 64, 256));
            if (more) gather_load(xr, P.obs, ids + ((it + 1) & 1) * R);  // next tile's rows, hidden behind L2 + head
            wait_done();
            for (int cp = 0; cp < 2; ++cp) {   // L2 epilogue: H2 = relu(D + b2)
                float va[2][32];
                tmem_ld32(T_WORK + t_lane + half * 128 + cp * 64, va[0]);
                tmem_ld32(T_WORK + t_lane + half * 128 + cp * 64 + 32, va[1]);
                tmem_wait_ld();
                if (rvalid) {
#pragma unroll
                    for (int cb = 0; cb < 2; ++cb) {
                        const int c0 = half * 128 + cp * 64 + cb * 32;
                        const float* v = va[cb];
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4)
                            *reinterpret_cast<float4*>(sH2 + img_off(row, c0 + j4 * 4, 256)) =
                                make_float4(to_tf32(fmaxf(v[j4 * 4] + bias[128 + c0 + j4 * 4], 0.f)), to_tf32(fmaxf(v[j4 * 4 + 1] + bias[128 + c0 + j4 * 4 + 1], 0.f)),
                                            to_tf32(fmaxf(v[j4 * 4 + 2] + bias[128 + c0 + j4 * 4 + 2], 0.f)), to_tf32(fmaxf(v[j4 * 4 + 3] + bias[128 + c0 + j4 * 4 + 3], 0.f)));
                    }
                }
            }
            RL_STAGE(stream_gemm(T_WORK, aH2, 256, 1, 256, 64, 16));
            wait_done();
            if (more) gather_store<false>(sX, xr);                       // sX is free since the L1 MMAs completed
            if (half == 0) {   // head epilogue + action rule, one row per valid lane
                float v[16];
                tmem_ld16(T_WORK + t_lane, v);
                tmem_wait_ld();
                const int i = tile * R + row;
                if (rvalid && i < total) {
                    float qv[8], ssum = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { qv[j] = v[j] + bias[384 + j]; ssum += qv[j]; }
                    const float val = v[8] + bias[384 + 8], mean = ssum * 0.125f;     // B = 1: per-row mean (PERD3QN.py:202)
#pragma unroll
                    for (int j = 0; j < 8; ++j) qv[j] = qv[j] + val - mean;
                    int best = 0;
#pragma unroll
                    for (int j = 1; j < 8; ++j) if (qv[j] > qv[best]) best = j;       // first maximum
                    const int rid = idc[row];
                    const int w = rid / S, slot = rid - w * S;
                    const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));
                    int a = best;                                                      // PERD3QN.py:204-210
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
            fence_before();
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(*tmem_slot, 256);
}

// ---- weight images (tf32-rounded) from the kernel-layout parameter buffer ----
__global__ void k_build_wimg_dueling(const float* __restrict__ p, float* __restrict__ wimg) {
    using L = Layout<RL_MODEL_DUELING>;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // W1: n = k1 (128), k = kx (160)
    if (i < 128 * 160) {
        const int n = i / 160, k = i - n * 160;
        wimg[(WI_W1 + k / 32) * CHUNK_F + img_off(n, k % 32, 32)] = to_tf32(p[L::OFF_W1T + k * 128 + n]);
    }
    // W2: W2t[k1][n2]
    if (i < 128 * 256) {
        const int k1 = i / 256, n2 = i - k1 * 256;
        const float v = to_tf32(p[L::OFF_W2T + i]);
        wimg[(WI_W2K + k1 / 16) * CHUNK_F + img_off(n2, k1 % 16, 16)] = v;      // B of L2: n = n2, k = k1
        wimg[(WI_W2T + n2 / 32) * CHUNK_F + img_off(k1, n2 % 32, 32)] = v;      // B of dH1: n = k1, k = n2
    }
    // head: Wh[n2][j], j < 9 (zero padded to 16)
    if (i < 256 * 16) {
        const int n2 = i / 16, j = i - n2 * 16;
        const float v = j < 9 ? to_tf32(p[L::OFF_WH + n2 * 9 + j]) : 0.f;
        wimg[WI_WH * CHUNK_F + img_off(j, n2, 256)] = v;                        // B of head: n = j, k = n2
        wimg[WI_WHT * CHUNK_F + img_off(n2, j, 16)] = v;                        // B of dH2: n = n2, k = j
    }
}

}  // namespace

extern "C" {

int rl_tc_wimg_floats(void) { return WI_CHUNKS_H * CHUNK_F; }      // (sized for the fp16 image, which carries 8 more chunks)

int rl_brain_build_wimg(int32_t kind, const float* params, float* wimg, void* stream) {
    RL_ARG_CHECK(params && wimg);
    if (kind != RL_MODEL_DUELING) return rl_set_err(RL_ERR_UNSUPPORTED, "tensor-core weight images: dueling networks only");
    k_build_wimg_dueling<<<(128 * 256 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(params, wimg);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_brain_build_wimg_h(int32_t kind, const float* params, void* wimg_h, void* stream) {
    RL_ARG_CHECK(params && wimg_h);
    if (kind != RL_MODEL_DUELING) return rl_set_err(RL_ERR_UNSUPPORTED, "tensor-core weight images: dueling networks only");
    k_build_wimg_dueling_h<<<(128 * 256 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(params, reinterpret_cast<__half*>(wimg_h));
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_brain_learn_h(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                     const int32_t* sample_idx, const rl_learn_bufs* learn, const void* wimg_eval_h, const void* wimg_target_h,
                     void* stream) {
    if (replay && replay->obs_fp16) return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_learn_h: float16 replay rows are read by rl_brain_learn_p / rl_brain_learn only");
    RL_ARG_CHECK(cfg && rows && replay && sample_idx && learn && wimg_eval_h && wimg_target_h);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1 && learn->batch == R);
    RL_ARG_CHECK(learn->params && learn->target && learn->grad_scratch && learn->grad && learn->new_prio && learn->loss);
    if (learn->kind != RL_MODEL_DUELING) return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_learn_h: dueling networks only");
    TcLearnParamsH PH;
    TcLearnParams& P = PH.p;
    P.cfg = *cfg;
    P.ev_rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_EVENT) * rows->row_cap;
    P.ev_total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    P.rp = *replay; P.sample_idx = sample_idx; P.lb = *learn; P.wimg_e = nullptr; P.wimg_t = nullptr; P.trace = nullptr;
    PH.wimg_e = reinterpret_cast<const __half*>(wimg_eval_h); PH.wimg_t = reinterpret_cast<const __half*>(wimg_target_h);
    static PerDeviceOnce attr;
    if (attr.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_learn_dueling_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCH_SMEM));
    }
    cudaStream_t st = (cudaStream_t)stream;
    k_learn_dueling_h<<<rl_learn_grid(), NTHREADS2, TCH_SMEM, st>>>(PH);
    RL_CUDA_CHECK(cudaGetLastError());
    return rl_learn_reduce(learn, P.ev_total, 1, (void*)st);
}

int rl_brain_act_tc(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                    const rl_brain_act* brain, const float* wimg_eval, uint64_t t_act, float* q_out, void* stream) {
    RL_ARG_CHECK(cfg && bufs && rows && brain && wimg_eval);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1);
    RL_ARG_CHECK(bufs->rec && bufs->obs_state && brain->params && brain->epsilon);
    if (brain->kind != RL_MODEL_DUELING || brain->rule != RL_ACT_DUELING)
        return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_act_tc: dueling networks only");
    TcActParams P;
    P.cfg = *cfg; P.rec = bufs->rec; P.obs = bufs->obs_state;
    P.rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_ALL) * rows->row_cap;
    P.total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_ALL;
    P.params = brain->params; P.wimg = wimg_eval; P.epsilon = brain->epsilon; P.t_act = t_act;
    P.q_out = q_out ? q_out + (size_t)gene * rows->row_cap * 8 : nullptr;
    static PerDeviceOnce attr;
    if (attr.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_act_dueling_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
    }
    k_act_dueling_tc<<<rl_learn_grid(), NTHREADS, TC_SMEM, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_brain_act_h(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                   const rl_brain_act* brain, const void* wimg_eval_h, uint64_t t_act, float* q_out, void* stream) {
    RL_ARG_CHECK(cfg && bufs && rows && brain && wimg_eval_h);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1);
    RL_ARG_CHECK(bufs->rec && bufs->obs_state && brain->params && brain->epsilon);
    if (brain->kind != RL_MODEL_DUELING || brain->rule != RL_ACT_DUELING)
        return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_act_h: dueling networks only");
    TcActParamsH PH;
    TcActParams& P = PH.p;
    P.cfg = *cfg; P.rec = bufs->rec; P.obs = bufs->obs_state;
    P.rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_ALL) * rows->row_cap;
    P.total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_ALL;
    P.params = brain->params; P.wimg = nullptr; P.epsilon = brain->epsilon; P.t_act = t_act;
    P.q_out = q_out ? q_out + (size_t)gene * rows->row_cap * 8 : nullptr;
    PH.wimg = reinterpret_cast<const __half*>(wimg_eval_h);
    static PerDeviceOnce attr;
    if (attr.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_act_dueling_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ACTH_SMEM));
    }
    k_act_dueling_h<<<rl_learn_grid(), NTHREADS2, ACTH_SMEM, (cudaStream_t)stream>>>(PH);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_brain_learn_tc(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                      const int32_t* sample_idx, const rl_learn_bufs* learn, const float* wimg_eval, const float* wimg_target,
                      void* stream) {
    if (replay && replay->obs_fp16) return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_learn_tc: float16 replay rows are read by rl_brain_learn_p / rl_brain_learn only");
    RL_ARG_CHECK(cfg && rows && replay && sample_idx && learn && wimg_eval && wimg_target);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1 && learn->batch == R);
    RL_ARG_CHECK(learn->params && learn->target && learn->grad_scratch && learn->grad && learn->new_prio && learn->loss);
    if (learn->kind != RL_MODEL_DUELING) return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_learn_tc: dueling networks only");
    TcLearnParams P;
    P.cfg = *cfg;
    P.ev_rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_EVENT) * rows->row_cap;
    P.ev_total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    P.rp = *replay; P.sample_idx = sample_idx; P.lb = *learn; P.wimg_e = wimg_eval; P.wimg_t = wimg_target;
    P.trace = nullptr;
    static long long* trace_dev = nullptr;
    const bool tracing = getenv("RL_TC_TRACE") != nullptr;
    if (tracing) {
        if (!trace_dev) RL_CUDA_CHECK(cudaMalloc(&trace_dev, sizeof(long long) * 8 * 40));
        RL_CUDA_CHECK(cudaMemsetAsync(trace_dev, 0, sizeof(long long) * 8 * 40, (cudaStream_t)stream));
        P.trace = trace_dev;
    }
    static PerDeviceOnce attr2;
    if (attr2.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_learn_dueling_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC2_SMEM));
    }
    const int n_cta = rl_learn_grid();
    cudaStream_t st = (cudaStream_t)stream;
    k_learn_dueling_tc2<<<n_cta, NTHREADS2, TC2_SMEM, st>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    if (tracing) {
        long long h[8 * 40];
        RL_CUDA_CHECK(cudaStreamSynchronize(st));
        RL_CUDA_CHECK(cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
        for (int it = 1; it < 4; ++it) {
            fprintf(stderr, "[tc trace] event %d (cycles since event start):", it);
            for (int k = 1; k < 30 && h[it * 40 + k]; ++k) fprintf(stderr, " %lld", h[it * 40 + k] - h[it * 40]);
            fprintf(stderr, " | issuer:");
            for (int k = 30; k < 40 && h[it * 40 + k]; ++k) fprintf(stderr, " %lld", h[it * 40 + k] - h[it * 40]);
            fprintf(stderr, "\n");
        }
    }
    return rl_learn_reduce(learn, P.ev_total, 1, (void*)st);
}

}  // extern "C"

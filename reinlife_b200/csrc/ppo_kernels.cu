// ppo_kernels.cu -- PPO training on the device (Models/PPO.py:62-77, 113-162).
//
// Reference semantics, per brain: every agent with age > 1 appends (s, a, r/100, s', pi_old(a), done) to ONE python list
// `data` shared by all agents of the gene (PPO.py:71-73, put_data :113); an agent whose age % train_freq == 0 or that
// died then calls learn() on whatever the list holds (:75-77), which consumes the list (make_batch :117-134) and runs
// k_epoch = 3 optimizer steps on it: td_target = r + gamma v(s') mask, delta = td_target - v(s), GAE as a reverse scan
// over the LIST ORDER (transitions of different agents interleaved -- a reference quirk that is kept), clipped
// surrogate + scalar smooth-L1 value loss, Adam.
//
// Here: one data list per (world, brain) in HBM (`traj`, rows appended in the reference's agent order).  A step's train
// triggers cut the list into SEGMENTS (one per trigger, in order); the rows left after the last trigger stay for the
// next step.  All segments of all worlds form one flat row list that 64-row tiles walk:
//   k_ppo_store   append + mark segment ends           k_ppo_scan / k_ppo_flat   flat row list, per-row segment length
//   per epoch: k_ppo_tiles<0> v(s), v(s') -> td, delta;  k_ppo_gae  reverse scan per segment;
//              k_ppo_tiles<1> pi(s), v(s) -> loss gradient -> explicit backward into per-CTA slabs;  reduce;  rl_brain_adam
//   k_ppo_compact  drop the consumed rows.
// N-world semantics as for the other brains: every segment is one learn() call against the same pre-step weights, the
// per-segment gradients (each a mean over its own T rows) are averaged, one Adam step per epoch.
#include <string.h>
#include "learn_tile.cuh"

namespace {

using namespace mlp;

struct PpoParams {
    rl_world_cfg cfg;
    rl_world_bufs wb;
    rl_rows_bufs rows;
    rl_ppo_bufs pb;
    rl_learn_bufs lb;
    const float* prob;          // pi_old(a) per row id of the act-time list (written by rl_brain_act_all)
    int32_t gene, train_freq;
};

constexpr int PT = 256;

// ---- append this step's transitions, mark the train triggers: CTA per world ----
__global__ void __launch_bounds__(PT) k_ppo_store(const PpoParams P) {
    __shared__ int s_last;
    const int w = blockIdx.x, NW = P.cfg.n_worlds, S = P.cfg.slot_cap, ld = P.cfg.obs_ld, TC = P.pb.traj.capacity;
    const int gk = P.gene * RL_N_ROW_KINDS + RL_ROWS_STORE;
    const int cnt = P.rows.count[(size_t)gk * NW + w];
    const int off = P.rows.offset[(size_t)gk * NW + w];
    const int len0 = P.pb.traj.len[w];
    if (threadIdx.x == 0) s_last = -1;
    __syncthreads();
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const rl_replay_bufs& tj = P.pb.traj;
    for (int tr = warp; tr < cnt; tr += PT / 32) {
        if (off + tr >= P.rows.row_cap) break;
        const int p = len0 + tr;
        if (p >= TC) { if (lane == 0 && P.pb.status) atomicOr(P.pb.status, 1); break; }
        const int row = P.rows.rows[(size_t)gk * P.rows.row_cap + off + tr];
        const int4 rv = reinterpret_cast<const int4*>(P.wb.rec)[row];
        const int prev = (rv.w >> 16) & 0xFFFF;
        const size_t q = (size_t)w * TC + p;
        const float4* s0 = reinterpret_cast<const float4*>(P.wb.obs_state + ((size_t)w * S + prev) * ld);
        const float4* s1 = reinterpret_cast<const float4*>(P.wb.obs_prime + (size_t)row * ld);
        float4* d0 = reinterpret_cast<float4*>(tj.obs + q * ld);
        float4* d1 = reinterpret_cast<float4*>(tj.next_obs + q * ld);
        for (int v = lane; v < ld / 4; v += 32) { d0[v] = __ldg(s0 + v); d1[v] = __ldg(s1 + v); }
        if (lane == 0) {
            const int age = (int16_t)(rv.y & 0xFFFF);
            const bool dead = (rv.w & RL_F_DEAD) != 0;
            tj.action[q] = (int8_t)((rv.w >> 8) & 0xFF);
            tj.reward[q] = P.wb.reward_div100[row];                     // reward / 100.0, PPO.py:73
            tj.done[q] = dead ? 1 : 0;
            tj.prio[q] = P.prob[(size_t)w * S + prev];                  // prob[action].item(), PPO.py:73
            const bool trig = (age % P.train_freq == 0) || dead;        // PPO.py:75
            P.pb.seg_end[q] = trig ? 1 : 0;
            if (trig) atomicMax(&s_last, p);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        P.pb.n_cons[w] = s_last + 1;                                    // rows consumed by this step's learn() calls
        tj.len[w] = min(TC, len0 + cnt);
    }
}

// ---- exclusive prefix of n_cons over the worlds of the shard: one CTA ----
__global__ void __launch_bounds__(1024) k_ppo_scan(const int32_t* __restrict__ n_cons, int32_t* __restrict__ row_off, int NW) {
    __shared__ int wsum[32];
    const int per = (NW + 1023) / 1024;
    const int i0 = min(NW, (int)threadIdx.x * per), i1 = min(NW, i0 + per);
    int local = 0;
    for (int i = i0; i < i1; ++i) local += n_cons[i];
    int incl = local;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int i = 0; i < warp; ++i) base += wsum[i];
    int run = base + incl - local;
    for (int i = i0; i < i1; ++i) { row_off[i] = run; run += n_cons[i]; }
    if (threadIdx.x == 1023) row_off[NW] = run;
}

// ---- flat row list + per-row segment length: CTA per world ----
__global__ void __launch_bounds__(PT) k_ppo_flat(const PpoParams P) {
    extern __shared__ __align__(16) unsigned char fs[];
    const int w = blockIdx.x, TC = P.pb.traj.capacity;
    const int n = P.pb.n_cons[w];
    if (n == 0) return;
    int32_t* sT = reinterpret_cast<int32_t*>(fs);
    uint8_t* sE = fs + (size_t)TC * 4;
    const uint8_t* ge = P.pb.seg_end + (size_t)w * TC;
    for (int j = threadIdx.x; j < n; j += PT) sE[j] = ge[j];
    __syncthreads();
    if (threadIdx.x == 0) {
        int start = 0;
        for (int j = 0; j < n; ++j)
            if (sE[j]) { const int T = j - start + 1; for (int k = start; k <= j; ++k) sT[k] = T; start = j + 1; }
    }
    __syncthreads();
    const int r0 = P.pb.row_off[w];
    for (int j = threadIdx.x; j < n; j += PT) {
        if (r0 + j >= P.pb.row_cap) { if (P.pb.status) atomicOr(P.pb.status, 2); break; }
        P.pb.flat_src[r0 + j] = w * TC + j;
        P.pb.row_T[r0 + j] = sT[j];
        P.pb.row_end[r0 + j] = sE[j];
    }
}

// ---- GAE (PPO.py:143-150): one thread per segment end walks its segment backwards, float32 arithmetic (NEP 50) ----
__global__ void k_ppo_gae(const PpoParams P, float gl) {
    const int NR = min(P.pb.row_off[P.cfg.n_worlds], P.pb.row_cap);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < NR; r += gridDim.x * blockDim.x) {
        if (!P.pb.row_end[r]) continue;
        const int T = P.pb.row_T[r];
        float run = 0.f;
        for (int k = 0; k < T; ++k) {
            run = __fadd_rn(__fmul_rn(gl, run), P.pb.delta[r - k]);
            P.pb.adv[r - k] = run;
        }
    }
}

// ---- 64-row tiles of the flat list.  MODE 0: v(s'), v(s) -> td_target, delta.  MODE 1: loss gradient + backward ----
constexpr int P_LDX = RL_K1 + 4, P_LDH = 256 + 4, P_WHN = 256 * 9 + 16;
constexpr size_t PPO_SMEM =
    sizeof(float) * ((size_t)R * P_LDX + 2 * (size_t)R * P_LDH + 2 * (CHUNK_BYTES / 4) + P_WHN + R * 16 + R * 12 + 2 * R) + sizeof(int) * R + 64;
static_assert(PPO_SMEM <= 227 * 1024, "shared memory budget");

template <int MODE>
__global__ void __launch_bounds__(NT, 1) k_ppo_tiles(const PpoParams P, float clip_lo, float clip_hi) {
    using M = Model<RL_MODEL_PPO>;
    using L = Layout<RL_MODEL_PPO>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* bufX = reinterpret_cast<float*>(smem_raw);
    float* bufH1 = bufX + (size_t)R * P_LDX;
    float* bufH2 = bufH1 + (size_t)R * P_LDH;
    float* wbuf = bufH2 + (size_t)R * P_LDH;
    float* Wh = wbuf + 2 * (CHUNK_BYTES / 4);
    float* outh = Wh + P_WHN;            // [64][16]
    float* dout = outh + R * 16;         // [64][12]
    float* vnext = dout + R * 12;        // [64]
    float* vcur = vnext + R;             // [64]
    int* idx = reinterpret_cast<int*>(vcur + R);
    uint64_t* bars = reinterpret_cast<uint64_t*>(idx + R);

    const float* Pw = P.lb.params;
    float* G = P.lb.grad_scratch + (size_t)blockIdx.x * L::N_TRAIN;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    for (int i = threadIdx.x; i < M::N2 * M::NH + M::NH; i += NT) Wh[i] = Pw[L::OFF_WH + i];
    if (MODE == 1)
        for (int i = threadIdx.x; i < L::N_TRAIN / 4; i += NT) reinterpret_cast<float4*>(G)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    Pipe pp{wbuf, bars, 0u};

    const int NR = min(P.pb.row_off[P.cfg.n_worlds], P.pb.row_cap);
    const int n_tiles = (NR + R - 1) / R;
    const rl_replay_bufs& tj = P.pb.traj;
    const float gamma = P.lb.gamma;
    // value head on the rows of bufH2: v[r] = bh[8] + sum_k H2[r][k] Wh[k][8]; 4 threads per row
    auto value_head = [&](float* out) {
        const int r = threadIdx.x >> 2, part = threadIdx.x & 3;
        const float* h = bufH2 + (size_t)r * P_LDH + part * 64;
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < 64; ++k) acc = fmaf(h[k], Wh[(part * 64 + k) * 9 + 8], acc);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (part == 0) out[r] = acc + Wh[M::N2 * 9 + 8];
        __syncthreads();
    };
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r_me = tile * R + (int)threadIdx.x;
        if (threadIdx.x < R) idx[threadIdx.x] = r_me < NR ? P.pb.flat_src[r_me] : 0;
        __syncthreads();
        if (MODE == 0) {
            gather64(bufX, P_LDX, tj.next_obs, idx);
            __syncthreads();
            gemm_stage<RL_K1, M::N1, 1, true>(bufX, P_LDX, Pw + L::OFF_W1T, Pw + L::OFF_B1, bufH1, P_LDH, pp);
            gemm_stage<M::N1, M::N2, 1, true>(bufH1, P_LDH, Pw + L::OFF_W2T, Pw + L::OFF_B2, bufH2, P_LDH, pp);
            value_head(vnext);
        }
        gather64(bufX, P_LDX, tj.obs, idx);
        __syncthreads();
        gemm_stage<RL_K1, M::N1, 1, true>(bufX, P_LDX, Pw + L::OFF_W1T, Pw + L::OFF_B1, bufH1, P_LDH, pp);
        gemm_stage<M::N1, M::N2, 1, true>(bufH1, P_LDH, Pw + L::OFF_W2T, Pw + L::OFF_B2, bufH2, P_LDH, pp);
        if (MODE == 0) {
            value_head(vcur);
            if (threadIdx.x < R && r_me < NR) {
                const int q = idx[threadIdx.x];
                const float mask = tj.done[q] ? 0.f : 1.f;                                        // PPO.py:126
                const float td = __fadd_rn(tj.reward[q], __fmul_rn(__fmul_rn(gamma, vnext[threadIdx.x]), mask));   // :140
                P.pb.td[r_me] = td;
                P.pb.delta[r_me] = __fsub_rn(td, vcur[threadIdx.x]);                              // :141
            }
            __syncthreads();
            continue;
        }
        // ---------------- MODE 1: pi(s), v(s), loss gradient (PPO.py:152-158) ----------------
        head64<M::N2, M::NH>(bufH2, P_LDH, Wh, outh);
        if (threadIdx.x < R) {
            const int b = threadIdx.x;
            float d[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) d[j] = 0.f;
            if (r_me < NR) {
                const int q = idx[b];
                const float* o = outh + b * 16;
                float mx = o[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) mx = fmaxf(mx, o[j]);
                float pi[8], s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { pi[j] = expf(o[j] - mx); s += pi[j]; }
#pragma unroll
                for (int j = 0; j < 8; ++j) pi[j] = pi[j] / s;
                const int a = tj.action[q] & 7;
                float pi_a = pi[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) pi_a = j == a ? pi[j] : pi_a;
                const float ratio = expf(logf(pi_a) - logf(tj.prio[q]));                           // :154
                const float A = P.pb.adv[r_me];
                const float surr1 = ratio * A;
                const float surr2 = fminf(fmaxf(ratio, clip_lo), clip_hi) * A;                     // :157
                const bool inside = ratio >= clip_lo && ratio <= clip_hi;
                const float invT = 1.0f / (float)P.pb.row_T[r_me];                                 // loss.mean() over the segment
                // d(-min(surr1, surr2))/d ratio: ties (ratio inside the clip range) split evenly between the two
                // branches and both reach `ratio`; outside the range only the unclipped branch carries gradient
                const float d_ratio = (inside || surr1 < surr2) ? -A * invT : 0.f;
                const float c = d_ratio * ratio;                                                   // d/d log pi_a
#pragma unroll
                for (int j = 0; j < 8; ++j) d[j] = c * ((j == a ? 1.f : 0.f) - pi[j]);
                const float dv = o[8] - P.pb.td[r_me];                                             // smooth_l1(v(s), td), mean over T
                d[8] = fminf(fmaxf(dv, -1.f), 1.f) * invT;
            }
#pragma unroll
            for (int j = 0; j < 9; ++j) dout[b * 12 + j] = d[j];
        }
        __syncthreads();
        // ---- head gradients: dWh[k][j] += sum_b H2[b][k] dOut[b][j], dbh[j] += sum_b dOut[b][j] ----
        {
            const int k = threadIdx.x;   // NT == N2
            float acc[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[j] = 0.f;
            for (int b = 0; b < R; ++b) {
                const float h = bufH2[(size_t)b * P_LDH + k];
#pragma unroll
                for (int j = 0; j < 9; ++j) acc[j] = fmaf(h, dout[b * 12 + j], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 9; ++j) G[L::OFF_WH + k * 9 + j] += acc[j];
            if (threadIdx.x < 9) {
                float s = 0.f;
                for (int b = 0; b < R; ++b) s += dout[b * 12 + threadIdx.x];
                G[L::OFF_BH + threadIdx.x] += s;
            }
        }
        __syncthreads();
        // ---- dH2 = (dOut Wh^T) * relu'(H2), in place ----
        for (int o = threadIdx.x; o < R * M::N2; o += NT) {
            const int b = o >> 8, k = o & 255;
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 9; ++j) v = fmaf(dout[b * 12 + j], Wh[k * 9 + j], v);
            float* h = bufH2 + (size_t)b * P_LDH + k;
            *h = *h > 0.f ? v : 0.f;
        }
        __syncthreads();
        // ---- dW2 += H1^T dH2 (two 128-row halves), db2 ----
        outer_accum<128, M::N2, 8, 16>(bufH1, P_LDH, bufH2, P_LDH, G + L::OFF_W2T);
        outer_accum<128, M::N2, 8, 16>(bufH1 + 128, P_LDH, bufH2, P_LDH, G + L::OFF_W2T + 128 * M::N2);
        colsum_accum<M::N2>(bufH2, P_LDH, G + L::OFF_B2);
        __syncthreads();
        // ---- dH1 = (dH2 W2) * relu'(H1), in place over H1 ----
        gemm_stage<M::N2, M::N1, 2, false>(bufH2, P_LDH, Pw + L::OFF_W2, nullptr, bufH1, P_LDH, pp);
        // ---- dW1 += X^T dH1 (two 128-column halves), db1 ----
        outer_accum<RL_K1, 128, 20, 4>(bufX, P_LDX, bufH1, P_LDH, G + L::OFF_W1T, M::N1);
        outer_accum<RL_K1, 128, 20, 4>(bufX, P_LDX, bufH1 + 128, P_LDH, G + L::OFF_W1T + 128, M::N1);
        colsum_accum<M::N1>(bufH1, P_LDH, G + L::OFF_B1);
        __syncthreads();
    }
}

// ---- drop the consumed rows: CTA per world, rows moved front-to-back in order (dst < src) ----
__global__ void __launch_bounds__(PT) k_ppo_compact(const PpoParams P) {
    const int w = blockIdx.x, ld = P.cfg.obs_ld, TC = P.pb.traj.capacity;
    const int n = P.pb.n_cons[w];
    if (n == 0) return;
    const rl_replay_bufs& tj = P.pb.traj;
    const int len = tj.len[w], left = max(0, len - n);
    const size_t base = (size_t)w * TC;
    for (int j = 0; j < left; ++j) {
        const size_t s = base + n + j, d = base + j;
        float4 a, b;
        const int v = threadIdx.x;
        const bool on = v < ld / 4;
        if (on) { a = reinterpret_cast<const float4*>(tj.obs + s * ld)[v]; b = reinterpret_cast<const float4*>(tj.next_obs + s * ld)[v]; }
        int8_t ac = 0; float rw = 0.f, pr = 0.f; uint8_t dn = 0;
        if (v == 0) { ac = tj.action[s]; rw = tj.reward[s]; pr = tj.prio[s]; dn = tj.done[s]; }
        __syncthreads();                                     // every read of row s happens before any write of row d <= s
        if (on) { reinterpret_cast<float4*>(tj.obs + d * ld)[v] = a; reinterpret_cast<float4*>(tj.next_obs + d * ld)[v] = b; }
        if (v == 0) { tj.action[d] = ac; tj.reward[d] = rw; tj.prio[d] = pr; tj.done[d] = dn; P.pb.seg_end[d] = 0; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { tj.len[w] = left; P.pb.n_cons[w] = 0; }
}

int ppo_fill(PpoParams& P, const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_ppo_bufs* ppo) {
    RL_ARG_CHECK(cfg && ppo);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1);
    const rl_replay_bufs& tj = ppo->traj;
    RL_ARG_CHECK(tj.obs && tj.next_obs && tj.action && tj.reward && tj.done && tj.prio && tj.len);
    RL_ARG_CHECK(tj.capacity > 0 && tj.capacity <= 8192 && (int64_t)cfg->n_worlds * tj.capacity < (1ll << 31));
    RL_ARG_CHECK(ppo->n_cons && ppo->row_off && ppo->seg_end && ppo->flat_src && ppo->row_T && ppo->row_end);
    RL_ARG_CHECK(ppo->td && ppo->delta && ppo->adv && ppo->row_cap > 0);
    memset(&P, 0, sizeof(P));
    P.cfg = *cfg;
    if (rows) P.rows = *rows;
    P.pb = *ppo; P.gene = gene; P.train_freq = 1;
    return RL_OK;
}

}  // namespace

extern "C" {

int rl_ppo_store(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                 const float* prob, int32_t train_freq, const rl_ppo_bufs* ppo, void* stream) {
    PpoParams P;
    RL_ARG_CHECK(bufs && rows && prob && train_freq > 0);
    int rc = ppo_fill(P, cfg, rows, gene, ppo);
    if (rc) return rc;
    RL_ARG_CHECK(bufs->rec && bufs->obs_state && bufs->obs_prime && bufs->reward_div100);
    P.wb = *bufs; P.prob = prob; P.train_freq = train_freq;
    cudaStream_t st = (cudaStream_t)stream;
    k_ppo_store<<<cfg->n_worlds, PT, 0, st>>>(P);
    k_ppo_scan<<<1, 1024, 0, st>>>(ppo->n_cons, ppo->row_off, cfg->n_worlds);
    k_ppo_flat<<<cfg->n_worlds, PT, (size_t)ppo->traj.capacity * 5, st>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_ppo_epoch(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_ppo_bufs* ppo,
                 const rl_learn_bufs* learn, void* stream) {
    PpoParams P;
    RL_ARG_CHECK(rows && learn);
    int rc = ppo_fill(P, cfg, rows, gene, ppo);
    if (rc) return rc;
    RL_ARG_CHECK(learn->kind == RL_MODEL_PPO && learn->params && learn->grad_scratch && learn->grad);
    P.lb = *learn;
    static PerDeviceOnce attr;
    if (attr.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_ppo_tiles<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PPO_SMEM));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_ppo_tiles<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PPO_SMEM));
    }
    cudaStream_t st = (cudaStream_t)stream;
    const float clip_lo = (float)(1.0 - (double)ppo->eps_clip), clip_hi = (float)(1.0 + (double)ppo->eps_clip);
    const float gl = (float)((double)learn->gamma * (double)ppo->lmbda);
    const int n_cta = rl_learn_grid();
    k_ppo_tiles<0><<<n_cta, NT, PPO_SMEM, st>>>(P, clip_lo, clip_hi);
    k_ppo_gae<<<2 * n_cta, 256, 0, st>>>(P, gl);
    k_ppo_tiles<1><<<n_cta, NT, PPO_SMEM, st>>>(P, clip_lo, clip_hi);
    RL_CUDA_CHECK(cudaGetLastError());
    return rl_learn_reduce(learn, rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_EVENT, 0, stream);
}

int rl_ppo_compact(const rl_world_cfg* cfg, int32_t gene, const rl_ppo_bufs* ppo, void* stream) {
    PpoParams P;
    int rc = ppo_fill(P, cfg, nullptr, gene, ppo);
    if (rc) return rc;
    k_ppo_compact<<<cfg->n_worlds, PT, 0, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

}  // extern "C"

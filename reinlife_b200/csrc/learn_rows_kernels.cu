// learn_rows_kernels.cu -- train steps whose loss is a sum over independent rows: DQN (Models/DQN.py:142-153) and
// PERDQN (Models/PERDQN.py:130-186; its 153-64-64-8 network runs in the DQN layout with the first hidden layer
// zero-padded from 64 to 128 units -- padded units stay exactly zero through forward, backward and Adam).
//
// Reference: train(q, q_target, memory, optimizer) runs 5 iterations of
//     s, a, r, s', done_mask = memory.sample(32);  q_a = q(s).gather(1, a);  y = r + gamma * max_a q_target(s') * done_mask
//     loss = F.smooth_l1_loss(q_a, y);  optimizer.zero_grad();  loss.backward();  optimizer.step()
// One launch = one of the 5 iterations for EVERY train event of the brain: rows never interact (no dueling mean), so a
// 64-row tile holds two 32-row events; rows of skipped events (sample_idx < 0: ring <= 1000 items, DQN.py:79) carry
// weight 0.  Same machinery as learn_kernels.cu: persistent CTA per SM, activations in shared memory, weights streamed
// by cp.async.bulk, per-CTA gradient slabs summed in a fixed order (deterministic).
#include "learn_tile.cuh"

namespace {

using namespace mlp;

struct RowsLearnParams {
    rl_world_cfg cfg;
    const int32_t* ev_rows;     // EVENT list of this brain
    const int32_t* ev_total;    // device scalar
    rl_replay_bufs rp;
    const int32_t* sample_idx;  // [row_cap, batch]
    const float* ev_weight;     // PERDQN: [row_cap] mean of the event's float32 importance weights (PERDQN.py:182)
    rl_learn_bufs lb;
};

// MODE 0: DQN  -- 32-row events, smooth-L1 (DQN.py:149).
// MODE 1: PERDQN -- 64-row events, loss = mean_i(is_w_i * mse_loss(pred, target)) with mse_loss a scalar mean
//         (PERDQN.py:182), errors |pred - target| written to lb.new_prio for Memory.update (PERDQN.py:168-174).
template <int MODE> struct RowsMode;
template <> struct RowsMode<0> { static constexpr int B = 32; };
template <> struct RowsMode<1> { static constexpr int B = 64; };

constexpr int DQ_LDX = RL_K1 + 4, DQ_LDH1 = 128 + 4, DQ_LDH2 = 64 + 4, DQ_WHN = 64 * 8 + 16;
constexpr size_t DQN_SMEM =
    sizeof(float) * ((size_t)R * DQ_LDX + (size_t)R * DQ_LDH1 + (size_t)R * DQ_LDH2 + 2 * (CHUNK_BYTES / 4) + 2 * DQ_WHN + R * 8 + R * 8 + 5 * R) +
    sizeof(int) * 2 * R + 64;

template <int MODE>
__global__ void __launch_bounds__(NT, 1) k_learn_dqn(const RowsLearnParams P) {
    constexpr int DQ_B = RowsMode<MODE>::B;
    using M = Model<RL_MODEL_DQN>;
    using L = Layout<RL_MODEL_DQN>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* bufX = reinterpret_cast<float*>(smem_raw);
    float* bufH1 = bufX + (size_t)R * DQ_LDX;
    float* bufH2 = bufH1 + (size_t)R * DQ_LDH1;
    float* wbuf = bufH2 + (size_t)R * DQ_LDH2;
    float* Wh_e = wbuf + 2 * (CHUNK_BYTES / 4);
    float* Wh_t = Wh_e + DQ_WHN;
    float* outh = Wh_t + DQ_WHN;         // [64][8]
    float* dout = outh + R * 8;          // [64][8]
    float* rew = dout + R * 8;
    float* dmask = rew + R;
    float* nq = dmask + R;
    float* wt = nq + R;
    float* lrow = wt + R;
    int* idx = reinterpret_cast<int*>(lrow + R);
    int* act = idx + R;
    uint64_t* bars = reinterpret_cast<uint64_t*>(act + R);

    const float* Pe = P.lb.params;
    const float* Pt = P.lb.target;
    float* G = P.lb.grad_scratch + (size_t)blockIdx.x * L::N_TRAIN;

    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    for (int i = threadIdx.x; i < M::N2 * M::NH + M::NH; i += NT) { Wh_e[i] = Pe[L::OFF_WH + i]; Wh_t[i] = Pt[L::OFF_WH + i]; }
    for (int i = threadIdx.x; i < L::N_TRAIN / 4; i += NT) reinterpret_cast<float4*>(G)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    Pipe pp{wbuf, bars, 0u};

    const int total = *P.ev_total;
    const int n_tiles = (total * DQ_B + R - 1) / R;
    const int S = P.cfg.slot_cap, cap = P.rp.capacity;
    const float gamma = P.lb.gamma;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x < R) {
            const int r = tile * R + threadIdx.x, e = r / DQ_B;
            int si = e < total ? P.sample_idx[r] : -1;
            const bool valid = si >= 0;
            const int gi = valid ? (P.ev_rows[e] / S) * cap + si : 0;     // transition index over all rings
            idx[threadIdx.x] = gi;
            act[threadIdx.x] = valid ? P.rp.action[gi] : 0;
            rew[threadIdx.x] = valid ? P.rp.reward[gi] : 0.f;
            dmask[threadIdx.x] = valid ? (P.rp.done[gi] ? 0.f : 1.f) : 0.f;   // done_mask, DQN.py:74-77
            wt[threadIdx.x] = valid ? 1.f : 0.f;
        }
        __syncthreads();
        // ---- q_target(s').max(1) (DQN.py:147) ----
        gather64(bufX, DQ_LDX, P.rp.next_obs, idx);
        __syncthreads();
        gemm_stage<RL_K1, M::N1, 1, true>(bufX, DQ_LDX, Pt + L::OFF_W1T, Pt + L::OFF_B1, bufH1, DQ_LDH1, pp);
        gemm_stage<M::N1, M::N2, 1, true>(bufH1, DQ_LDH1, Pt + L::OFF_W2T, Pt + L::OFF_B2, bufH2, DQ_LDH2, pp);
        for (int o = threadIdx.x; o < R * 8; o += NT) {
            const int r = o >> 3, j = o & 7;
            const float* h = bufH2 + (size_t)r * DQ_LDH2;
            float acc = Wh_t[M::N2 * 8 + j];
#pragma unroll 8
            for (int k = 0; k < M::N2; ++k) acc = fmaf(h[k], Wh_t[k * 8 + j], acc);
            outh[o] = acc;
        }
        __syncthreads();
        if (threadIdx.x < R) {
            const float* o = outh + threadIdx.x * 8;
            float mx = o[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) mx = fmaxf(mx, o[j]);
            nq[threadIdx.x] = mx;
        }
        __syncthreads();
        // ---- q(s) (DQN.py:145), activations kept for the backward ----
        gather64(bufX, DQ_LDX, P.rp.obs, idx);
        __syncthreads();
        gemm_stage<RL_K1, M::N1, 1, true>(bufX, DQ_LDX, Pe + L::OFF_W1T, Pe + L::OFF_B1, bufH1, DQ_LDH1, pp);
        gemm_stage<M::N1, M::N2, 1, true>(bufH1, DQ_LDH1, Pe + L::OFF_W2T, Pe + L::OFF_B2, bufH2, DQ_LDH2, pp);
        for (int o = threadIdx.x; o < R * 8; o += NT) {
            const int r = o >> 3, j = o & 7;
            const float* h = bufH2 + (size_t)r * DQ_LDH2;
            float acc = Wh_e[M::N2 * 8 + j];
#pragma unroll 8
            for (int k = 0; k < M::N2; ++k) acc = fmaf(h[k], Wh_e[k * 8 + j], acc);
            outh[o] = acc;
        }
        __syncthreads();
        // ---- loss and its gradient w.r.t. q(s)[a] ----
        if (threadIdx.x < R) {
            const int b = threadIdx.x;
            const float qa = outh[b * 8 + act[b]];
            const float y = rew[b] + gamma * nq[b] * dmask[b];                       // DQN.py:148, PERDQN.py:163
            const float d = qa - y, ad = fabsf(d);
            const int e = (tile * R + b) / DQ_B;
            float l, g;
            if (MODE == 0) {                                                         // smooth-L1, beta = 1, mean over 32 rows
                l = (ad < 1.f ? 0.5f * d * d : ad - 0.5f) * wt[b];
                g = fminf(fmaxf(d, -1.f), 1.f) * (1.0f / DQ_B) * wt[b];
            } else {                                                                 // mean(is_w) * mse
                const float ew = e < total ? P.ev_weight[e] : 0.f;
                l = d * d * wt[b] * ew;
                g = ew * 2.f * d * (1.0f / DQ_B) * wt[b];
                if (e < total) P.lb.new_prio[(size_t)tile * R + b] = ad;             // errors, PERDQN.py:166
            }
            lrow[b] = l;
#pragma unroll
            for (int j = 0; j < 8; ++j) dout[b * 8 + j] = j == act[b] ? g : 0.f;
        }
        __syncthreads();
        if (threadIdx.x < R / DQ_B) {                                                // per-event loss, fixed summation order
            const int e = tile * (R / DQ_B) + threadIdx.x;
            float l = 0.f;
            for (int b = 0; b < DQ_B; ++b) l += lrow[threadIdx.x * DQ_B + b];
            if (e < total) P.lb.loss[e] = l * (1.0f / DQ_B);
        }
        // ---- fc3 gradients: dWh[k][j] += sum_b H2[b][k] dOut[b][j]; dbh ----
        {
            const int k = threadIdx.x >> 2, j0 = (threadIdx.x & 3) * 2;
            float a0 = 0.f, a1 = 0.f;
            for (int b = 0; b < R; ++b) {
                const float h = bufH2[(size_t)b * DQ_LDH2 + k];
                a0 = fmaf(h, dout[b * 8 + j0], a0); a1 = fmaf(h, dout[b * 8 + j0 + 1], a1);
            }
            G[L::OFF_WH + k * 8 + j0] += a0; G[L::OFF_WH + k * 8 + j0 + 1] += a1;
            if (threadIdx.x < 8) {
                float s = 0.f;
                for (int b = 0; b < R; ++b) s += dout[b * 8 + threadIdx.x];
                G[L::OFF_BH + threadIdx.x] += s;
            }
        }
        __syncthreads();
        // ---- dH2 = (dOut Wh^T) * relu'(H2), in place ----
        for (int o = threadIdx.x; o < R * M::N2; o += NT) {
            const int b = o / M::N2, k = o - b * M::N2;
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) v = fmaf(dout[b * 8 + j], Wh_e[k * 8 + j], v);
            float* h = bufH2 + (size_t)b * DQ_LDH2 + k;
            *h = *h > 0.f ? v : 0.f;
        }
        __syncthreads();
        outer_accum<M::N1, M::N2, 8, 4>(bufH1, DQ_LDH1, bufH2, DQ_LDH2, G + L::OFF_W2T);
        colsum_accum<M::N2>(bufH2, DQ_LDH2, G + L::OFF_B2);
        __syncthreads();
        gemm_stage<M::N2, M::N1, 2, false>(bufH2, DQ_LDH2, Pe + L::OFF_W2, nullptr, bufH1, DQ_LDH1, pp);
        outer_accum<RL_K1, M::N1, 20, 4>(bufX, DQ_LDX, bufH1, DQ_LDH1, G + L::OFF_W1T);
        colsum_accum<M::N1>(bufH1, DQ_LDH1, G + L::OFF_B1);
        __syncthreads();
    }
}

// grad[n_train] = number of events that were not skipped by the sampler
__global__ void k_count_valid(const int32_t* __restrict__ sample_idx, int batch, const int32_t* ev_total, float* out) {
    __shared__ int red[8];
    const int total = *ev_total;
    int c = 0;
    for (int e = threadIdx.x; e < total; e += blockDim.x) c += sample_idx[(size_t)e * batch] >= 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int i = 0; i < (int)blockDim.x / 32; ++i) s += red[i];
        *out = (float)s;
    }
}

}  // namespace

// (shared with tc_dqn_kernels.cu)
int rl_count_valid_launch(const int32_t* sample_idx, int batch, const int32_t* ev_total, float* out, void* stream) {
    k_count_valid<<<1, 256, 0, (cudaStream_t)stream>>>(sample_idx, batch, ev_total, out);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

extern "C" {

int rl_brain_learn_dqn(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                       const int32_t* sample_idx, const rl_learn_bufs* learn, void* stream) {
    if (replay && replay->obs_fp16) return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_learn_dqn: float16 replay rows are not supported");
    RL_ARG_CHECK(cfg && rows && replay && sample_idx && learn);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1);
    RL_ARG_CHECK(learn->kind == RL_MODEL_DQN && learn->batch == 32);
    RL_ARG_CHECK(learn->params && learn->target && learn->grad_scratch && learn->grad && learn->loss);
    RL_ARG_CHECK((int64_t)cfg->n_worlds * replay->capacity < (1ll << 31));
    RowsLearnParams P;
    P.cfg = *cfg;
    P.ev_rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_EVENT) * rows->row_cap;
    P.ev_total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    P.rp = *replay; P.sample_idx = sample_idx; P.ev_weight = nullptr; P.lb = *learn;
    static PerDeviceOnce attr;
    if (attr.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_learn_dqn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DQN_SMEM));
    }
    cudaStream_t st = (cudaStream_t)stream;
    k_learn_dqn<0><<<rl_learn_grid(), NT, DQN_SMEM, st>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    int rc = rl_learn_reduce(learn, P.ev_total, 0, stream);
    if (rc) return rc;
    k_count_valid<<<1, 256, 0, st>>>(sample_idx, 32, P.ev_total, learn->grad + Layout<RL_MODEL_DQN>::N_TRAIN);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_brain_learn_perdqn(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                          const int32_t* sample_idx, const float* ev_weight, const rl_learn_bufs* learn, void* stream) {
    if (replay && replay->obs_fp16) return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_learn_perdqn: float16 replay rows are not supported");
    RL_ARG_CHECK(cfg && rows && replay && sample_idx && ev_weight && learn);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes && cfg->obs_ld == RL_K1);
    RL_ARG_CHECK(learn->kind == RL_MODEL_DQN && learn->batch == 64);
    RL_ARG_CHECK(learn->params && learn->target && learn->grad_scratch && learn->grad && learn->loss && learn->new_prio);
    RL_ARG_CHECK((int64_t)cfg->n_worlds * replay->capacity < (1ll << 31));
    RowsLearnParams P;
    P.cfg = *cfg;
    P.ev_rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_EVENT) * rows->row_cap;
    P.ev_total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    P.rp = *replay; P.sample_idx = sample_idx; P.ev_weight = ev_weight; P.lb = *learn;
    static PerDeviceOnce attr;
    if (attr.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_learn_dqn<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DQN_SMEM));
    }
    cudaStream_t st = (cudaStream_t)stream;
    k_learn_dqn<1><<<rl_learn_grid(), NT, DQN_SMEM, st>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    int rc = rl_learn_reduce(learn, P.ev_total, 0, stream);
    if (rc) return rc;
    k_count_valid<<<1, 256, 0, st>>>(sample_idx, 64, P.ev_total, learn->grad + Layout<RL_MODEL_DQN>::N_TRAIN);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

}  // extern "C"

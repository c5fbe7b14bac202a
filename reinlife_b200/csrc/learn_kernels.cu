// learn_kernels.cu -- batched train() events for the dueling brains (PERD3QN / D3QN), Adam, target sync.
//
// Reference: PERD3QNAgent.train (Models/PERD3QN.py:94-115) == D3QNAgent.train (Models/D3QN.py:97-116) up to the
// sampler: q = eval(obs); q' = target(next_obs); y = r + gamma*(1-done)*max_a q'; MSE(q[a], y); priorities
// |max_a q' - q[a]|; backward; torch.optim.Adam(lr).  The dueling combine uses the mean of the WHOLE [64,8]
// advantage tensor (PERD3QN.py:202), which couples the rows of an event in forward and backward
//  -- so one event = one 64-row tile = one CTA iteration.
//
// One persistent CTA per SM walks the brain's EVENT list.  Activations never leave shared memory; the per-event
// weight gradients (54k floats) are accumulated into a CTA-private, L2-resident scratch slab and summed across
// CTAs in a fixed order afterwards (deterministic, no atomics).
#include "learn_tile.cuh"

namespace {

using namespace mlp;

struct LearnParams {
    rl_world_cfg cfg;
    const int32_t* ev_rows;     // EVENT list of this brain
    const int32_t* ev_total;    // device scalar
    rl_replay_bufs rp;
    const int32_t* sample_idx;  // [row_cap, 64]
    rl_learn_bufs lb;
    int32_t n_cta;
};

constexpr int LDX = RL_K1 + 4, LDH1 = 128 + 4, LDH2 = 256 + 4, WHN = 256 * 9 + 16;
constexpr size_t LEARN_SMEM =
    sizeof(float) * ((size_t)R * LDX + (size_t)R * LDH1 + (size_t)R * LDH2 + 2 * (CHUNK_BYTES / 4) + 2 * WHN + R * 16 + R * 12 + 4 * R + 32) +
    sizeof(int) * 2 * R + 64;

__global__ void __launch_bounds__(NT, 1) k_learn_dueling(const LearnParams P) {
    using M = Model<RL_MODEL_DUELING>;
    using L = Layout<RL_MODEL_DUELING>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* bufX = reinterpret_cast<float*>(smem_raw);
    float* bufH1 = bufX + (size_t)R * LDX;
    float* bufH2 = bufH1 + (size_t)R * LDH1;
    float* wbuf = bufH2 + (size_t)R * LDH2;
    float* Wh_e = wbuf + 2 * (CHUNK_BYTES / 4);
    float* Wh_t = Wh_e + WHN;
    float* outh = Wh_t + WHN;            // [64][16]
    float* dout = outh + R * 16;         // [64][12]
    float* rew = dout + R * 12;
    float* dn = rew + R;
    float* nq = dn + R;
    float* gb = nq + R;
    float* red = gb + R;                 // [32]
    int* idx = reinterpret_cast<int*>(red + 32);
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
    const int S = P.cfg.slot_cap, cap = P.rp.capacity;
    const float gamma = P.lb.gamma;
    for (int e = blockIdx.x; e < total; e += gridDim.x) {
        const int w = P.ev_rows[e] / S;
        const size_t ring = (size_t)w * cap;
        if (P.sample_idx[(size_t)e * R] < 0) continue;     // event skipped by the uniform sampler (ring shorter than a batch)
        if (threadIdx.x < R) {
            const int i = P.sample_idx[(size_t)e * R + threadIdx.x];
            idx[threadIdx.x] = i;
            act[threadIdx.x] = P.rp.action[ring + i];
            rew[threadIdx.x] = P.rp.reward[ring + i];
            dn[threadIdx.x] = (float)P.rp.done[ring + i];
        }
        __syncthreads();
        // ---- target forward on next_obs (PERD3QN.py:104-105) ----
        if (P.rp.obs_fp16) gather64_h(bufX, LDX, reinterpret_cast<const __half*>(P.rp.next_obs) + ring * RL_K1, idx);
        else gather64(bufX, LDX, P.rp.next_obs + ring * RL_K1, idx);
        __syncthreads();
        gemm_stage<RL_K1, M::N1, 1, true>(bufX, LDX, Pt + L::OFF_W1T, Pt + L::OFF_B1, bufH1, LDH1, pp);
        gemm_stage<M::N1, M::N2, 1, true>(bufH1, LDH1, Pt + L::OFF_W2T, Pt + L::OFF_B2, bufH2, LDH2, pp);
        head64<M::N2, M::NH>(bufH2, LDH2, Wh_t, outh);
        {
            float s = 0.f;
            for (int o = threadIdx.x; o < R * 8; o += NT) s += outh[(o >> 3) * 16 + (o & 7)];
            const float mean_t = block_sum(s, red) * (1.0f / (8 * R));
            if (threadIdx.x < R) {
                const float* o = outh + threadIdx.x * 16;
                float mx = o[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) mx = fmaxf(mx, o[j]);
                nq[threadIdx.x] = mx + o[8] - mean_t;
            }
        }
        __syncthreads();
        // ---- eval forward on obs (:103), activations kept for the backward ----
        if (P.rp.obs_fp16) gather64_h(bufX, LDX, reinterpret_cast<const __half*>(P.rp.obs) + ring * RL_K1, idx);
        else gather64(bufX, LDX, P.rp.obs + ring * RL_K1, idx);
        __syncthreads();
        gemm_stage<RL_K1, M::N1, 1, true>(bufX, LDX, Pe + L::OFF_W1T, Pe + L::OFF_B1, bufH1, LDH1, pp);
        gemm_stage<M::N1, M::N2, 1, true>(bufH1, LDH1, Pe + L::OFF_W2T, Pe + L::OFF_B2, bufH2, LDH2, pp);
        head64<M::N2, M::NH>(bufH2, LDH2, Wh_e, outh);
        {
            float s = 0.f;
            for (int o = threadIdx.x; o < R * 8; o += NT) s += outh[(o >> 3) * 16 + (o & 7)];
            const float mean_e = block_sum(s, red) * (1.0f / (8 * R));
            float g = 0.f, sq = 0.f;
            if (threadIdx.x < R) {
                const int b = threadIdx.x;
                const float qa = outh[b * 16 + act[b]] + outh[b * 16 + 8] - mean_e;       // :106
                const float y = rew[b] + gamma * (1.0f - dn[b]) * nq[b];                  // :107
                const float diff = qa - y;
                g = 2.0f * diff * (1.0f / R);                                            // d MSE / d q_a
                sq = diff * diff;
                gb[b] = g;
                P.lb.new_prio[(size_t)e * R + b] = fabsf(nq[b] - qa);                     // :110
            }
            const float gsum = block_sum(g, red);
            const float loss = block_sum(sq, red) * (1.0f / R);
            if (threadIdx.x == 0) P.lb.loss[e] = loss;
            const float shift = gsum * (1.0f / (8 * R));     // every advantage entry gets -sum_b g_b / (8B)
            for (int o = threadIdx.x; o < R * 9; o += NT) {
                const int b = o / 9, j = o - b * 9;
                dout[b * 12 + j] = j == 8 ? gb[b] : ((j == act[b] ? gb[b] : 0.f) - shift);
            }
        }
        __syncthreads();
        // ---- head gradients: dWh[k][j] = sum_b H2[b][k] dOut[b][j], dbh[j] = sum_b dOut[b][j] ----
        {
            const int k = threadIdx.x;   // NT == N2
            float acc[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[j] = 0.f;
            for (int b = 0; b < R; ++b) {
                const float h = bufH2[(size_t)b * LDH2 + k];
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
            for (int j = 0; j < 9; ++j) v = fmaf(dout[b * 12 + j], Wh_e[k * 9 + j], v);
            float* h = bufH2 + (size_t)b * LDH2 + k;
            *h = *h > 0.f ? v : 0.f;
        }
        __syncthreads();
        // ---- dW2 += H1^T dH2, db2 += colsum(dH2) ----
        outer_accum<M::N1, M::N2, 8, 16>(bufH1, LDH1, bufH2, LDH2, G + L::OFF_W2T);
        colsum_accum<M::N2>(bufH2, LDH2, G + L::OFF_B2);
        __syncthreads();
        // ---- dH1 = (dH2 W2) * relu'(H1), in place over H1 (W2 output-major copy = k-major for this product) ----
        gemm_stage<M::N2, M::N1, 2, false>(bufH2, LDH2, Pe + L::OFF_W2, nullptr, bufH1, LDH1, pp);
        // ---- dW1 += X^T dH1, db1 += colsum(dH1) ----
        outer_accum<RL_K1, M::N1, 20, 4>(bufX, LDX, bufH1, LDH1, G + L::OFF_W1T);
        colsum_accum<M::N1>(bufH1, LDH1, G + L::OFF_B1);
        __syncthreads();
    }
}

// grad[p] = sum over CTA slabs (fixed order); grad[n_train] = number of events
__global__ void k_grad_reduce(const float* __restrict__ scratch, int n_cta, int n_train, const int32_t* ev_total,
                              float* __restrict__ grad, int w1_rowmajor, int n1) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_train) {
        int src = p;
        if (w1_rowmajor && p < RL_K1 * n1) {      // W1 slab layouts of the tensor-core kernels: 1 = [k1][kx], 2 = [kx / 4][k1][kx % 4]
            const int kx = p / n1, k1 = p - kx * n1;
            src = w1_rowmajor == 2 ? ((kx >> 2) * n1 + k1) * 4 + (kx & 3) : k1 * RL_K1 + kx;
        }
        float s = 0.f;
        for (int c = 0; c < n_cta; ++c) s += scratch[(size_t)c * n_train + src];
        grad[p] = s;
    }
    if (p == 0) grad[n_train] = (float)(*ev_total);
}

// torch.optim.Adam, defaults (betas .9/.999, eps 1e-8, no weight decay / amsgrad), single-tensor formulas
template <int KIND>
__global__ void k_adam(const rl_learn_bufs lb) {
    using M = Model<KIND>; using L = Layout<KIND>;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= L::N_TRAIN) return;
    const float cnt = lb.grad[L::N_TRAIN];
    if (!(cnt > 0.f)) return;
    const int t = *lb.adam_step + 1;
    const float g = (lb.grad[p] / cnt) * lb.mask[p];
    const double bc1 = 1.0 - pow(0.9, (double)t), bc2 = 1.0 - pow(0.999, (double)t);
    const float step_size = (float)((double)lb.lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    float m = lb.adam_m[p], v = lb.adam_v[p], w = lb.params[p];
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), 0.1f));                                   // exp_avg.lerp_(grad, 1-beta1)
    v = __fadd_rn(__fmul_rn(v, 0.999f), __fmul_rn(__fmul_rn(0.001f, g), g));              // mul_(beta2).addcmul_(g, g, 1-beta2)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), 1e-8f);
    w = __fsub_rn(w, __fmul_rn(step_size, __fdiv_rn(m, denom)));                          // addcdiv_(m, denom, -step_size)
    lb.adam_m[p] = m; lb.adam_v[p] = v; lb.params[p] = w;
    if (p >= L::OFF_W2T && p < L::OFF_B2) {
        const int q = p - L::OFF_W2T, k = q / M::N2, n = q - k * M::N2;
        lb.params[L::OFF_W2 + n * M::N1 + k] = w;
    }
}

__global__ void k_adam_tick(const rl_learn_bufs lb, int n_train) {
    if (lb.grad[n_train] > 0.f) *lb.adam_step += 1;
}

int n_train_of(int kind) {
    if (kind == RL_MODEL_DUELING) return Layout<RL_MODEL_DUELING>::N_TRAIN;
    if (kind == RL_MODEL_DQN) return Layout<RL_MODEL_DQN>::N_TRAIN;
    return Layout<RL_MODEL_PPO>::N_TRAIN;
}
int n_total_of(int kind) {
    if (kind == RL_MODEL_DUELING) return Layout<RL_MODEL_DUELING>::N_TOTAL;
    if (kind == RL_MODEL_DQN) return Layout<RL_MODEL_DQN>::N_TOTAL;
    return Layout<RL_MODEL_PPO>::N_TOTAL;
}

__global__ void k_sync_target(const float* __restrict__ src, float* __restrict__ dst, int n, const int32_t* cond) {
    if (cond && *cond <= 0) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace

int rl_learn_reduce(const rl_learn_bufs* learn, const int32_t* ev_total, int w1_rowmajor, void* stream) {
    const int nt = n_train_of(learn->kind);
    const int n1 = learn->kind == RL_MODEL_PPO ? 256 : 128;
    k_grad_reduce<<<(nt + 255) / 256, 256, 0, (cudaStream_t)stream>>>(learn->grad_scratch, rl_learn_grid(), nt, ev_total, learn->grad,
                                                                      w1_rowmajor, n1);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

extern "C" {

int rl_learn_grid(void) { return rl_device_sm_count(); }

int rl_brain_learn(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                   const int32_t* sample_idx, const rl_learn_bufs* learn, void* stream) {
    RL_ARG_CHECK(cfg && rows && replay && sample_idx && learn);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes);
    RL_ARG_CHECK(learn->params && learn->target && learn->grad_scratch && learn->grad && learn->new_prio && learn->loss);
    RL_ARG_CHECK(cfg->obs_ld == RL_K1);
    if (learn->kind != RL_MODEL_DUELING)
        return rl_set_err(RL_ERR_UNSUPPORTED, "rl_brain_learn: model kind %d not implemented yet (dueling only)", learn->kind);
    RL_ARG_CHECK(learn->batch == R);
    LearnParams P;
    P.cfg = *cfg;
    P.ev_rows = rows->rows + (size_t)(gene * RL_N_ROW_KINDS + RL_ROWS_EVENT) * rows->row_cap;
    P.ev_total = rows->total + gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    P.rp = *replay; P.sample_idx = sample_idx; P.lb = *learn; P.n_cta = rl_learn_grid();
    static PerDeviceOnce attr;
    if (attr.need()) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_learn_dueling, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LEARN_SMEM));
    }
    cudaStream_t st = (cudaStream_t)stream;
    k_learn_dueling<<<P.n_cta, NT, LEARN_SMEM, st>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return rl_learn_reduce(learn, P.ev_total, 0, stream);
}

int rl_brain_adam(const rl_learn_bufs* learn, void* stream) {
    RL_ARG_CHECK(learn && learn->params && learn->grad && learn->adam_m && learn->adam_v && learn->mask && learn->adam_step);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = n_train_of(learn->kind);
    const int blocks = (nt + 255) / 256;
    if (learn->kind == RL_MODEL_DUELING) k_adam<RL_MODEL_DUELING><<<blocks, 256, 0, st>>>(*learn);
    else if (learn->kind == RL_MODEL_DQN) k_adam<RL_MODEL_DQN><<<blocks, 256, 0, st>>>(*learn);
    else if (learn->kind == RL_MODEL_PPO) k_adam<RL_MODEL_PPO><<<blocks, 256, 0, st>>>(*learn);
    else return rl_set_err(RL_ERR_ARG, "unknown model kind %d", learn->kind);
    k_adam_tick<<<1, 1, 0, st>>>(*learn, nt);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_brain_sync_target(const rl_learn_bufs* learn, const int32_t* cond, void* stream) {
    RL_ARG_CHECK(learn && learn->params && learn->target);
    k_sync_target<<<64, 256, 0, (cudaStream_t)stream>>>(learn->params, learn->target, n_total_of(learn->kind), cond);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

}  // extern "C"

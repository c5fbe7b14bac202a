// brain_kernels.cu -- batched brain.get_action for every listed agent of every world.
//
// Reference: the per-agent B=1 forwards of Helpers/trainer.py:88-89 / Helpers/tester.py:58-68 through
// Models/PERD3QN.py:81-89,198-210, Models/D3QN.py:82-93,161-173, Models/DQN.py:65-78,126-139,
// Models/PPO.py:54-60,101-106,164-169, Models/PERDQN.py:101-111,311-323 (DQN layout, first hidden layer zero-padded).  Here: persistent CTAs walk 64-row tiles of the per-brain row
// list; the whole 3-stage network runs out of shared memory (mlp_tile.cuh), and the exploration rule +
// argmax are fused into the epilogue that writes rec[].action.
#include "mlp_tile.cuh"
#include "models.cuh"

namespace {

using namespace mlp;

struct ActParams {
    rl_world_cfg cfg;
    rl_agent_rec* rec;
    const float* obs;          // obs_state
    const int32_t* rows;       // row list of this brain, kind ALL
    const int32_t* total;      // device scalar
    const float* params;
    const double* epsilon;
    uint64_t t_act;
    int32_t rule;
    float* q_out;              // [row_cap][8] or null
    float* prob_out;           // [n_worlds*slot_cap] or null
};

template <int KIND> constexpr int act_lda() { return (Model<KIND>::N2 > RL_K1 ? Model<KIND>::N2 : RL_K1) + 4; }
template <int KIND> constexpr int act_ldb() { return Model<KIND>::N1 + 4; }
template <int KIND> constexpr size_t act_smem() {
    return sizeof(float) * ((size_t)R * act_lda<KIND>() + (size_t)R * act_ldb<KIND>() + 2 * (CHUNK_BYTES / 4) +
                            (size_t)Model<KIND>::N2 * Model<KIND>::NH + 16) + 64;
}

// shared device pieces -------------------------------------------------------------------------------

// gather 64 observation rows (160 floats each) into shared memory; rows past `nrows` are zero
__device__ __forceinline__ void gather_rows(float* dst, int ldd, const float* __restrict__ src, int ld_src,
                                            const int32_t* __restrict__ ids, int nrows) {
    for (int v = threadIdx.x; v < R * (RL_K1 / 4); v += NT) {
        const int r = v / (RL_K1 / 4), c4 = v - r * (RL_K1 / 4);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows) x = __ldg(reinterpret_cast<const float4*>(src + (size_t)ids[r] * ld_src) + c4);
        *reinterpret_cast<float4*>(dst + (size_t)r * ldd + c4 * 4) = x;
    }
}

// head: OUT[r][j] = bh[j] + sum_k H2[r][k] * Wh[k][j]   (Wh, bh in shared memory)
template <int N2, int NH>
__device__ __forceinline__ void head_stage(const float* H2, int ldh, const float* Wh_s, float* OUT, int ldo) {
    for (int o = threadIdx.x; o < R * NH; o += NT) {
        const int r = o / NH, j = o - r * NH;
        const float* h = H2 + (size_t)r * ldh;
        float acc = Wh_s[N2 * NH + j];
#pragma unroll 8
        for (int k = 0; k < N2; ++k) acc = fmaf(h[k], Wh_s[k * NH + j], acc);
        OUT[(size_t)r * ldo + j] = acc;
    }
    __syncthreads();
}

template <int KIND>
__global__ void __launch_bounds__(NT, 1) k_brain_act(const ActParams P) {
    using M = Model<KIND>;
    using L = Layout<KIND>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int LDA = act_lda<KIND>(), LDB = act_ldb<KIND>();
    float* bufA = reinterpret_cast<float*>(smem_raw);
    float* bufB = bufA + (size_t)R * LDA;
    float* wbuf = bufB + (size_t)R * LDB;
    float* Wh_s = wbuf + 2 * (CHUNK_BYTES / 4);
    uint64_t* bars = reinterpret_cast<uint64_t*>(Wh_s + M::N2 * M::NH + 16);
    __shared__ int32_t ids[R];

    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    for (int i = threadIdx.x; i < M::N2 * M::NH + M::NH; i += NT) Wh_s[i] = P.params[L::OFF_WH + i];
    __syncthreads();
    Pipe pp{wbuf, bars, 0u};

    const int total = *P.total;
    const int S = P.cfg.slot_cap;
    const double epsilon = *P.epsilon;
    for (int tile = blockIdx.x; tile * R < total; tile += gridDim.x) {
        const int nrows = min(R, total - tile * R);
        if (threadIdx.x < R) ids[threadIdx.x] = threadIdx.x < nrows ? P.rows[tile * R + threadIdx.x] : 0;
        __syncthreads();
        gather_rows(bufA, LDA, P.obs, P.cfg.obs_ld, ids, nrows);
        __syncthreads();
        gemm_stage<RL_K1, M::N1, 1, true>(bufA, LDA, P.params + L::OFF_W1T, P.params + L::OFF_B1, bufB, LDB, pp);
        gemm_stage<M::N1, M::N2, 1, true>(bufB, LDB, P.params + L::OFF_W2T, P.params + L::OFF_B2, bufA, LDA, pp);
        head_stage<M::N2, M::NH>(bufA, LDA, Wh_s, bufB, 16);

        if (threadIdx.x < nrows) {
            const int r = threadIdx.x;
            const float* o = bufB + r * 16;
            float q[8];
            if (KIND == RL_MODEL_DUELING) {           // Q = A + V - mean(A), B=1 so the mean is per row (PERD3QN.py:202)
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) s += o[j];
                const float mean = s * 0.125f;
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = o[j] + o[8] - mean;
            } else if (KIND == RL_MODEL_PPO) {        // softmax(fc_pi), PPO.py:101-106
                float mx = o[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) mx = fmaxf(mx, o[j]);
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { q[j] = expf(o[j] - mx); s += q[j]; }
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = q[j] / s;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = o[j];
            }
            int best = 0;
#pragma unroll
            for (int j = 1; j < 8; ++j) if (q[j] > q[best]) best = j;     // first maximum
            const int row = ids[r];
            const int w = row / S, slot = row - w * S;
            const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));
            int a = best;
            if (P.rule == RL_ACT_DUELING) {           // PERD3QN.py:204-210
                const double u = rl_uniform(rl_draw(key, P.t_act, RL_SITE_ACT_EXPLORE, (uint32_t)slot));
                if (!(u > epsilon)) a = (int)rl_below(rl_draw(key, P.t_act, RL_SITE_ACT_RANDOM, (uint32_t)slot), 8);
            } else if (P.rule == RL_ACT_DQN) {        // DQN.py:135-139
                const double coin = rl_uniform(rl_draw(key, P.t_act, RL_SITE_ACT_EXPLORE, (uint32_t)slot));
                if (coin < epsilon) a = (int)rl_below(rl_draw(key, P.t_act, RL_SITE_ACT_RANDOM, (uint32_t)slot), 8);
            } else if (P.rule == RL_ACT_PERDQN) {     // PERDQN.py:101-111: np.random.rand() <= eps -> random.randrange(8)
                const double u = rl_uniform(rl_draw(key, P.t_act, RL_SITE_ACT_EXPLORE, (uint32_t)slot));
                if (u <= epsilon) a = (int)rl_below(rl_draw(key, P.t_act, RL_SITE_ACT_RANDOM, (uint32_t)slot), 8);
            } else {                                   // PPO.py:164-169: categorical by inverse CDF on one uniform
                const double u = rl_uniform(rl_draw(key, P.t_act, RL_SITE_ACT_SAMPLE, (uint32_t)slot));
                float c = 0.f;
                a = 7;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c += q[j]; if (a == 7 && u < (double)c) a = j; }
                if (P.prob_out) P.prob_out[row] = q[a];
            }
            reinterpret_cast<int8_t*>(P.rec + row)[13] = (int8_t)a;
            if (P.q_out) {
                float4* qo = reinterpret_cast<float4*>(P.q_out + (size_t)(tile * R + r) * 8);
                qo[0] = make_float4(q[0], q[1], q[2], q[3]);
                qo[1] = make_float4(q[4], q[5], q[6], q[7]);
            }
        }
        __syncthreads();
    }
}

struct EpsParams {
    rl_brain_sched sched[RL_MAX_GENES];
    const int32_t* total;
    double* eps;
    int64_t* seen;
    int64_t n_epi;
    int32_t n_brains;
};

__global__ void k_epsilon_update(const EpsParams P) {
    const int g = threadIdx.x;
    if (g >= P.n_brains) return;
    if (P.total[g * RL_N_ROW_KINDS + RL_ROWS_ALL] <= 0) return;      // get_action was not called for this brain
    const rl_brain_sched s = P.sched[g];
    if (!s.training) return;
    if (s.rule == RL_ACT_DUELING) {
        if (P.n_epi > P.seen[g]) {
            if (P.eps[g] > s.eps_min) P.eps[g] = P.eps[g] * s.decay;
            P.seen[g] = P.n_epi;
        }
    } else if (s.rule == RL_ACT_DQN) {
        if (P.n_epi % 30 == 0) P.eps[g] = fmax(0.01, 0.20 - 0.20 * ((double)P.n_epi / (double)s.max_epi));
    }
}

int sm_count() { return rl_device_sm_count(); }

template <int KIND>
int launch_act(const ActParams& P, cudaStream_t st) {
    static bool attr_set = false;
    constexpr size_t smem = act_smem<KIND>();
    if (!attr_set) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_brain_act<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    k_brain_act<KIND><<<sm_count(), NT, smem, st>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

template <int KIND> void fill_dims(rl_model_dims* d) {
    using M = Model<KIND>; using L = Layout<KIND>;
    d->n1 = M::N1; d->n2 = M::N2; d->nh = M::NH;
    d->off_b1 = L::OFF_B1; d->off_w2t = L::OFF_W2T; d->off_b2 = L::OFF_B2; d->off_wh = L::OFF_WH; d->off_bh = L::OFF_BH;
    d->off_w2 = L::OFF_W2; d->n_train = L::N_TRAIN; d->n_total = L::N_TOTAL;
}

}  // namespace

extern "C" {

int rl_model_get_dims(int32_t kind, rl_model_dims* out) {
    RL_ARG_CHECK(out);
    if (kind == RL_MODEL_DUELING) fill_dims<RL_MODEL_DUELING>(out);
    else if (kind == RL_MODEL_DQN) fill_dims<RL_MODEL_DQN>(out);
    else if (kind == RL_MODEL_PPO) fill_dims<RL_MODEL_PPO>(out);
    else return rl_set_err(RL_ERR_ARG, "unknown model kind %d", kind);
    return RL_OK;
}

int rl_brain_epsilon_update(const rl_rows_bufs* rows, const rl_brain_sched* sched, int32_t n_brains, int64_t n_epi,
                            double* eps_dev, int64_t* seen_dev, void* stream) {
    RL_ARG_CHECK(rows && sched && eps_dev && seen_dev && n_brains > 0 && n_brains <= RL_MAX_GENES);
    EpsParams P;
    for (int g = 0; g < n_brains; ++g) P.sched[g] = sched[g];
    P.total = rows->total; P.eps = eps_dev; P.seen = seen_dev; P.n_epi = n_epi; P.n_brains = n_brains;
    k_epsilon_update<<<1, 32, 0, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_brain_act_all(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows,
                     const rl_brain_act* brains, int32_t n_brains, uint64_t t_act, float* q_out, float* prob_out,
                     void* stream) {
    RL_ARG_CHECK(cfg && bufs && rows && brains && n_brains == cfg->n_genes);
    RL_ARG_CHECK(cfg->obs_ld == RL_K1);
    for (int g = 0; g < n_brains; ++g) {
        if (brains[g].kind < 0) continue;            // handled elsewhere (rl_brain_act_tc)
        ActParams P;
        P.cfg = *cfg; P.rec = bufs->rec; P.obs = bufs->obs_state;
        P.rows = rows->rows + (size_t)(g * RL_N_ROW_KINDS + RL_ROWS_ALL) * rows->row_cap;
        P.total = rows->total + g * RL_N_ROW_KINDS + RL_ROWS_ALL;
        P.params = brains[g].params; P.epsilon = brains[g].epsilon;
        RL_ARG_CHECK(P.epsilon != nullptr); P.t_act = t_act; P.rule = brains[g].rule;
        P.q_out = q_out ? q_out + (size_t)g * rows->row_cap * 8 : nullptr;
        P.prob_out = prob_out;
        RL_ARG_CHECK(P.params != nullptr);
        int rc;
        if (brains[g].kind == RL_MODEL_DUELING) rc = launch_act<RL_MODEL_DUELING>(P, (cudaStream_t)stream);
        else if (brains[g].kind == RL_MODEL_DQN) rc = launch_act<RL_MODEL_DQN>(P, (cudaStream_t)stream);
        else if (brains[g].kind == RL_MODEL_PPO) rc = launch_act<RL_MODEL_PPO>(P, (cudaStream_t)stream);
        else return rl_set_err(RL_ERR_ARG, "unknown model kind %d", brains[g].kind);
        if (rc) return rc;
    }
    return RL_OK;
}

}  // extern "C"

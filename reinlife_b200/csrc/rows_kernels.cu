// rows_kernels.cu -- deterministic per-brain compaction of the agent lists of all worlds.
// count (warp per world) -> scan over worlds (CTA per gene x kind) -> scatter (warp per world).
// Replaces the reference's per-agent Python loops `for agent in env.agents: agent.get_action / agent.learn`
// (Helpers/trainer.py:88-96) with dense per-brain row lists the batched brain kernels consume.
#include "rl_common.cuh"

namespace {

struct RowsParams {
    rl_world_cfg cfg;
    const rl_agent_rec* rec;
    const int32_t* n_agents;
    rl_rows_bufs r;
    int32_t train_freq[RL_MAX_GENES];
    int32_t event_on[RL_MAX_GENES];
    int32_t kinds_mask;
};

__device__ __forceinline__ bool row_pred(int kind, int age, unsigned flags, int train_freq, int event_on) {
    if (kind == RL_ROWS_ALL) return true;
    if (age <= 1) return false;                                        // entities.py:196
    if (kind == RL_ROWS_STORE) return true;
    return event_on && ((age % train_freq) == 0 || (flags & RL_F_DEAD));   // PERD3QN.py:120-122, DQN.py:88, PPO.py:75
}

template <bool SCATTER>
__global__ void __launch_bounds__(256) k_rows_pass(const RowsParams P) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    if (warp >= P.cfg.n_worlds) return;
    const int w = warp, G = P.cfg.n_genes, S = P.cfg.slot_cap, NW = P.cfg.n_worlds;
    const int n = min(P.n_agents[w], S);
    const int4* rg = reinterpret_cast<const int4*>(P.rec + (size_t)w * S);
    for (int g = 0; g < G; ++g) {
        const int tf = max(1, P.train_freq[g]), on = P.event_on[g];
        int run[RL_N_ROW_KINDS] = {0, 0, 0};
        int base[RL_N_ROW_KINDS] = {0, 0, 0};
        if (SCATTER)
            for (int k = 0; k < RL_N_ROW_KINDS; ++k)
                if (P.kinds_mask >> k & 1) base[k] = P.r.offset[(size_t)(g * RL_N_ROW_KINDS + k) * NW + w];
        for (int s0 = 0; s0 < n; s0 += 32) {
            const int sl = s0 + lane;
            bool mine = false; int age = 0; unsigned fl = 0;
            if (sl < n) {
                int4 v = rg[sl];
                mine = v.z == g; age = (int16_t)(v.y & 0xFFFF); fl = v.w & 0xFF;
            }
#pragma unroll
            for (int k = 0; k < RL_N_ROW_KINDS; ++k) {
                if (!(P.kinds_mask >> k & 1)) continue;
                const bool p = mine && row_pred(k, age, fl, tf, on);
                const unsigned m = __ballot_sync(0xffffffffu, p);
                if (SCATTER && p) {
                    const int pos = base[k] + run[k] + __popc(m & lanemask_lt());
                    if (pos < P.r.row_cap) P.r.rows[(size_t)(g * RL_N_ROW_KINDS + k) * P.r.row_cap + pos] = w * S + sl;
                }
                run[k] += __popc(m);
            }
        }
        if (!SCATTER && lane == 0)
            for (int k = 0; k < RL_N_ROW_KINDS; ++k)
                if (P.kinds_mask >> k & 1) P.r.count[(size_t)(g * RL_N_ROW_KINDS + k) * NW + w] = run[k];
    }
}

// exclusive scan over worlds, one CTA per (gene, kind)
__global__ void __launch_bounds__(1024) k_rows_scan(const RowsParams P) {
    const int gk = blockIdx.x;
    if (!(P.kinds_mask >> (gk % RL_N_ROW_KINDS) & 1)) return;
    const int NW = P.cfg.n_worlds;
    const int32_t* cnt = P.r.count + (size_t)gk * NW;
    int32_t* off = P.r.offset + (size_t)gk * NW;
    __shared__ int wsum[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    for (int base = 0; base < NW; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < NW ? cnt[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int x = wsum[lane], xi = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, xi, o);
                if (lane >= o) xi += t;
            }
            wsum[lane] = xi - x;   // exclusive
        }
        __syncthreads();
        const int carry = carry_s;
        if (i < NW) off[i] = carry + wsum[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wsum[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) P.r.total[gk] = min(carry_s, P.r.row_cap);
}

}  // namespace

extern "C" int rl_rows_build(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows,
                             const int32_t* train_freq_host, const int32_t* event_on_host, int32_t kinds_mask,
                             void* stream) {
    RL_ARG_CHECK(cfg && bufs && rows && rows->count && rows->offset && rows->total && rows->rows);
    RL_ARG_CHECK(cfg->n_genes > 0 && cfg->n_genes <= RL_MAX_GENES && rows->row_cap > 0);
    RowsParams P;
    P.cfg = *cfg; P.rec = bufs->rec; P.n_agents = bufs->n_agents; P.r = *rows; P.kinds_mask = kinds_mask;
    for (int g = 0; g < RL_MAX_GENES; ++g) {
        P.train_freq[g] = (train_freq_host && g < cfg->n_genes) ? train_freq_host[g] : 1;
        P.event_on[g] = (event_on_host && g < cfg->n_genes) ? event_on_host[g] : 0;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (cfg->n_worlds * 32 + 255) / 256;
    k_rows_pass<false><<<blocks, 256, 0, st>>>(P);
    k_rows_scan<<<cfg->n_genes * RL_N_ROW_KINDS, 1024, 0, st>>>(P);
    k_rows_pass<true><<<blocks, 256, 0, st>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

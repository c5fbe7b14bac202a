// stats_kernels.cu -- per-gene tracker partial sums over all worlds (Helpers/tracker.py:178-266), deterministic.
#include "rl_common.cuh"

namespace {

constexpr int ST = 256;
constexpr int MAXB = 148 * 4;      // one to two worlds per warp at 4096 worlds (a warp walks its worlds serially: latency-bound)

struct StatsParams {
    rl_world_cfg cfg;
    const rl_agent_rec* rec;
    const int32_t* n_agents;
    const float* reward;
    const int64_t* ctrl;
    double* scratch;      // [nblocks][G*8 + 8]
    int32_t* counter;
    double* out;          // [G*8 + 8]
    int32_t nblocks, per_block;
};

__global__ void __launch_bounds__(ST) k_world_stats(const StatsParams P) {
    const int G = P.cfg.n_genes, S = P.cfg.slot_cap, NV = G * RL_N_STATS + 8;
    __shared__ double acc[RL_MAX_GENES * RL_N_STATS + 8];
    __shared__ double wrew[ST / 32][RL_MAX_GENES];      // per-warp reward partials (merged in warp order: deterministic)
    __shared__ int is_last;
    for (int i = threadIdx.x; i < NV; i += ST) acc[i] = 0.0;
    for (int i = threadIdx.x; i < (ST / 32) * RL_MAX_GENES; i += ST) (&wrew[0][0])[i] = 0.0;
    __syncthreads();
    const int w0 = blockIdx.x * P.per_block, w1 = min(P.cfg.n_worlds, w0 + P.per_block);
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    // one warp per world, fixed world->warp assignment.  Counts / age sums / attacks / kills are integers (exact in
    // double, so the shared-memory atomics are order-independent); REWARD_SUM is not: every warp adds its worlds'
    // sums into its own shared slot in world order, and the slots are merged in warp order below.
    for (int w = w0 + warp; w < w1; w += ST / 32) {
        const int n = min(P.n_agents[w], S);
        unsigned present = 0;
        for (int g = 0; g < G; ++g) {
            double cnt = 0, age = 0, rew = 0, amax = 0, att = 0, kil = 0;
            for (int s0 = 0; s0 < n; s0 += 32) {
                const int sl = s0 + lane;
                if (sl < n) {
                    const int4 v = reinterpret_cast<const int4*>(P.rec + (size_t)w * S)[sl];
                    if (v.z == g) {
                        const int a = (int16_t)(v.y & 0xFFFF);
                        cnt += 1; age += a; amax = fmax(amax, (double)a);
                        rew += (double)P.reward[(size_t)w * S + sl];
                        att += ((int8_t)((v.w >> 8) & 0xFF)) >= 4 ? 1 : 0;
                        kil += (v.w & RL_F_KILLED) ? 1 : 0;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                cnt += __shfl_xor_sync(0xffffffffu, cnt, o); age += __shfl_xor_sync(0xffffffffu, age, o);
                rew += __shfl_xor_sync(0xffffffffu, rew, o); att += __shfl_xor_sync(0xffffffffu, att, o);
                kil += __shfl_xor_sync(0xffffffffu, kil, o); amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            }
            if (lane == 0 && cnt > 0) {
                double* a = acc + g * RL_N_STATS;
                atomicAdd(&a[RL_STAT_COUNT], cnt); atomicAdd(&a[RL_STAT_AGE_SUM], age);
                wrew[warp][g] += rew;
                atomicAdd(&a[RL_STAT_ATTACKS], att); atomicAdd(&a[RL_STAT_KILLS], kil); atomicAdd(&a[6], 1.0);
                // max via CAS on the bit pattern (non-negative doubles order like integers)
                atomicMax(reinterpret_cast<unsigned long long*>(&a[RL_STAT_AGE_MAX]), (unsigned long long)__double_as_longlong(amax));
            }
            if (cnt > 0) present |= 1u << g;
        }
        if (lane == 0) {
            double* t = acc + G * RL_N_STATS;
            atomicAdd(&t[0], (double)n);
            if (n > 0) atomicAdd(&t[1], 1.0);
            atomicAdd(&t[2], (double)__popc(present));
        }
    }
    __syncthreads();
    if (threadIdx.x < G) {
        double r = 0.0;
        for (int k = 0; k < ST / 32; ++k) r += wrew[k][threadIdx.x];
        acc[threadIdx.x * RL_N_STATS + RL_STAT_REWARD_SUM] = r;
    }
    __syncthreads();
    double* mine = P.scratch + (size_t)blockIdx.x * NV;
    for (int i = threadIdx.x; i < NV; i += ST) mine[i] = acc[i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(P.counter, 1) == P.nblocks - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // one warp per value: lane l folds blocks l, l + 32, ... in order, then a fixed xor tree -- deterministic, and the loads of a lane are
    // independent (a single thread walking all blocks of a value was a chain of ~nblocks / 4 L2 round trips: most of the kernel's time)
    for (int i = warp; i < NV; i += ST / 32) {
        const bool is_max = i < G * RL_N_STATS && (i % RL_N_STATS) == RL_STAT_AGE_MAX;
        double s = 0.0;
        for (int b = lane; b < P.nblocks; b += 32) {
            const double v = __ldcg(P.scratch + (size_t)b * NV + i);
            s = is_max ? fmax(s, v) : s + v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double t = __shfl_xor_sync(0xffffffffu, s, o);
            s = is_max ? fmax(s, t) : s + t;
        }
        if (lane == 0) P.out[i] = s;
    }
    if (threadIdx.x == 0) {
        P.out[G * RL_N_STATS + 3] = P.ctrl ? (double)P.ctrl[0] : 0.0;
        *P.counter = 0;
    }
}

int blocks_for(const rl_world_cfg* cfg) { return cfg->n_worlds < MAXB * 8 ? (cfg->n_worlds + 7) / 8 : MAXB; }

}  // namespace

extern "C" {

int rl_world_stats_scratch_doubles(const rl_world_cfg* cfg) {
    if (!cfg) return 0;
    return blocks_for(cfg) * (cfg->n_genes * RL_N_STATS + 8);
}

int rl_world_stats(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const int64_t* ctrl_dev, double* scratch_dev,
                   int32_t* counter_dev, double* out_dev, void* stream) {
    RL_ARG_CHECK(cfg && bufs && scratch_dev && counter_dev && out_dev);
    RL_ARG_CHECK(cfg->n_genes > 0 && cfg->n_genes <= RL_MAX_GENES);
    StatsParams P;
    P.cfg = *cfg; P.rec = bufs->rec; P.n_agents = bufs->n_agents; P.reward = bufs->reward; P.ctrl = ctrl_dev;
    P.scratch = scratch_dev; P.counter = counter_dev; P.out = out_dev;
    P.nblocks = blocks_for(cfg);
    P.per_block = (cfg->n_worlds + P.nblocks - 1) / P.nblocks;
    k_world_stats<<<P.nblocks, ST, 0, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

}  // extern "C"

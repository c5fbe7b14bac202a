// sumtree_kernels.cu -- PERDQN's prioritized memory (Models/PERDQN.py:198-308) on the device, one tree per (world, brain).
//
// Reference: SumTree = implicit binary heap of 2*capacity-1 float64 nodes (leaf of data slot d = d + capacity - 1);
// `update(idx, p)` writes the leaf and adds `change = p - old` to every ancestor (incremental, so rounding depends on
// the order of the calls); Memory.add stores new items with (|error| + 0.01)^0.6, Memory.sample draws one
// random.uniform per stratum of total/64, redrawing while the leaf holds no data, and returns importance weights
// (n * p/total)^-beta / max; Memory.update re-prioritises the 64 sampled leaves one by one.
//
// Two quirks of the reference are reproduced exactly (pinned by tests/golden/brain_golden3.npz):
//  * append_sample's `old_val` aliases the tensor it is compared with (PERDQN.py:119-126): the stored error is always 0,
//    so new items enter with the constant priority `p_new` = float32(0.01)^0.6 -- the two B=1 forwards of append_sample
//    have no effect on any result and are not evaluated;
//  * on the add path `p` is a float32 torch scalar, which turns `p - tree[idx]` and `tree[parent] += change` into
//    FLOAT32 tensor arithmetic (node rounded to float32, result stored back as float64); on the update path `p` is a
//    numpy float32 scalar and the same lines run in float64.
//
// Propagation is sequential per tree (one warp per world walks its adds / updates in reference order) and parallel
// across levels: lane l owns the ancestor l+1 levels above the leaf, every lane adds the same `change` to its own
// node, so each node sees exactly the reference's sequence of additions.
#include "rl_common.cuh"

namespace {

constexpr int ST = 256;

struct TreeParams {
    rl_world_cfg cfg;
    rl_rows_bufs rows;
    rl_replay_bufs rp;
    rl_sumtree_bufs tr;
    int32_t gene, batch;
    uint64_t t;
    int32_t* sample_idx;     // [row_cap, batch] data slots (-1: event skipped)
    float* ev_weight;        // [row_cap]
    const float* errors;     // [row_cap, batch]
};

// one SumTree.update(leaf, p): lane 0 owns the leaf, lane l >= 1 the ancestor l levels up (heap index h >> l, 1-based)
template <bool F32>
__device__ __forceinline__ void tree_update(double* tree, int leaf, float p, int lane) {
    double change_d = 0.0;
    float change_f = 0.f;
    if (lane == 0) {
        const double old = tree[leaf];
        if (F32) change_f = __fsub_rn(p, (float)old); else change_d = (double)p - old;
        tree[leaf] = (double)p;
    }
    change_d = __shfl_sync(0xffffffffu, change_d, 0);
    change_f = __shfl_sync(0xffffffffu, change_f, 0);
    const unsigned h = (unsigned)leaf + 1u;
    if (lane >= 1) {
        const unsigned a = h >> lane;
        if (a >= 1u) {
            double* node = tree + (a - 1u);
            if (F32) *node = (double)__fadd_rn((float)*node, change_f); else *node = *node + change_d;
        }
    }
    __syncwarp();
}

// ---- Memory.add for every STORE row of the gene, in agent order (before rl_replay_store moves pos/len) ----
__global__ void __launch_bounds__(ST) k_sumtree_add(const TreeParams P) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= P.cfg.n_worlds) return;
    const int NW = P.cfg.n_worlds, cap = P.tr.capacity, lane = lane_id();
    const int gk = P.gene * RL_N_ROW_KINDS + RL_ROWS_STORE;
    int cnt = P.rows.count[(size_t)gk * NW + w];
    if (cnt == 0) return;
    const int off = P.rows.offset[(size_t)gk * NW + w];
    cnt = min(cnt, max(0, P.rows.row_cap - off));
    double* tree = P.tr.tree + (size_t)w * (2 * cap - 1);
    int write = P.rp.pos[w];
    for (int i = 0; i < cnt; ++i) {
        tree_update<true>(tree, write + cap - 1, P.tr.p_new, lane);
        write = write + 1 >= cap ? 0 : write + 1;
    }
}

// ---- Memory.sample(batch) for every EVENT row: 64 threads per world, events of a world one after the other ----
__device__ __forceinline__ float priority_of(float err) {         // Memory._get_priority on a numpy float32 scalar (powf)
    return (float)pow((double)__fadd_rn(fabsf(err), 0.01f), (double)0.6f);
}

__global__ void __launch_bounds__(64) k_sumtree_sample(const TreeParams P) {
    __shared__ double red[64];
    __shared__ float redf[64];
    const int w = blockIdx.x, NW = P.cfg.n_worlds, cap = P.tr.capacity, batch = P.batch, i = threadIdx.x;
    const int gk = P.gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    int cnt = P.rows.count[(size_t)gk * NW + w];
    if (cnt == 0) return;
    const int off = P.rows.offset[(size_t)gk * NW + w];
    cnt = min(cnt, max(0, P.rows.row_cap - off));
    const int n_entries = P.rp.len[w];
    const double* tree = P.tr.tree + (size_t)w * (2 * cap - 1);
    const int n_nodes = 2 * cap - 1;
    const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));
    double beta = P.tr.beta[w];
    for (int e = 0; e < cnt; ++e) {
        const size_t q = (size_t)(off + e) * batch + i;
        if (n_entries < P.tr.train_start || n_entries <= 0) {            // learn(): train_model only once n_entries >= train_start
            if (i < batch) P.sample_idx[q] = -1;
            if (i == 0) P.ev_weight[off + e] = 0.f;
            continue;
        }
        beta = fmin(1.0, beta + 0.001);                                   // PERDQN.py:284
        const double total = tree[0];
        const double segment = total / (double)batch;
        const double a = segment * (double)i, b = segment * (double)(i + 1);
        int idx = 0, slot = 0;
        bool ok = false;
        for (int tries = 0; tries < 64 && !ok; ++tries) {                 // `while True` of PERDQN.py:290-295, bounded
            const double u = rl_uniform(rl_draw(key, P.t, RL_SITE_SUMTREE_SAMPLE, (uint32_t)((e * batch + i) * 64 + tries)));
            double s = __dadd_rn(a, __dmul_rn(b - a, u));                 // random.uniform(a, b); no fused multiply-add
            idx = 0;
            for (;;) {                                                    // SumTree._retrieve
                const int left = 2 * idx + 1;
                if (left >= n_nodes) break;
                const double tl = tree[left];
                if (s <= tl) idx = left; else { s -= tl; idx = left + 1; }
            }
            slot = idx - cap + 1;
            ok = slot < n_entries;
        }
        if (!ok) { slot = 0; idx = cap - 1; if (P.tr.status) atomicOr(P.tr.status, 1); }
        const double prob = tree[idx] / total;                           // PERDQN.py:300-302
        double wgt = pow((double)n_entries * prob, -beta);
        red[i] = wgt;
        __syncthreads();
        double mx = red[0];
        for (int j = 1; j < batch; ++j) mx = fmax(mx, red[j]);
        redf[i] = (float)(wgt / mx);                                      // torch.FloatTensor(is_weights)
        __syncthreads();
        if (i == 0) {
            float s = 0.f;
            for (int j = 0; j < batch; ++j) s += redf[j];
            P.ev_weight[off + e] = s / (float)batch;
        }
        P.sample_idx[q] = slot;
        __syncthreads();
    }
    if (i == 0) P.tr.beta[w] = beta;
}

// ---- Memory.update for the sampled leaves of every trained event, event by event, batch order, duplicates included ----
__global__ void __launch_bounds__(ST) k_sumtree_update(const TreeParams P) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= P.cfg.n_worlds) return;
    const int NW = P.cfg.n_worlds, cap = P.tr.capacity, batch = P.batch, lane = lane_id();
    const int gk = P.gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    int cnt = P.rows.count[(size_t)gk * NW + w];
    if (cnt == 0) return;
    const int off = P.rows.offset[(size_t)gk * NW + w];
    cnt = min(cnt, max(0, P.rows.row_cap - off));
    double* tree = P.tr.tree + (size_t)w * (2 * cap - 1);
    for (int e = 0; e < cnt; ++e) {
        for (int i = 0; i < batch; ++i) {
            const size_t q = (size_t)(off + e) * batch + i;
            const int slot = P.sample_idx[q];
            if (slot < 0) break;                                          // skipped event
            tree_update<false>(tree, slot + cap - 1, priority_of(P.errors[q]), lane);
        }
    }
}

// train_model's epsilon step (PERDQN.py:132-133), once per optimizer step that happened
__global__ void k_perdqn_eps(const float* grad_count, double* eps, double eps_min, double eps_decay) {
    if (*grad_count > 0.f && *eps > eps_min) *eps -= eps_decay;
}

int fill(TreeParams& P, const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* rp,
         const rl_sumtree_bufs* tr) {
    RL_ARG_CHECK(cfg && rows && rp && tr);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes);
    RL_ARG_CHECK(tr->tree && tr->beta && tr->capacity > 0 && tr->capacity == rp->capacity && rp->len && rp->pos);
    RL_ARG_CHECK(tr->capacity <= (1 << 30));                 // depth <= 31: one lane per ancestor level
    P.cfg = *cfg; P.rows = *rows; P.rp = *rp; P.tr = *tr; P.gene = gene; P.batch = 0; P.t = 0;
    P.sample_idx = nullptr; P.ev_weight = nullptr; P.errors = nullptr;
    return RL_OK;
}

}  // namespace

extern "C" {

int rl_sumtree_add(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                   const rl_sumtree_bufs* tree, void* stream) {
    TreeParams P;
    int rc = fill(P, cfg, rows, gene, replay, tree);
    if (rc) return rc;
    k_sumtree_add<<<(cfg->n_worlds * 32 + ST - 1) / ST, ST, 0, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_sumtree_sample(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                      const rl_sumtree_bufs* tree, int32_t batch, uint64_t t, int32_t* sample_idx, float* ev_weight,
                      void* stream) {
    TreeParams P;
    int rc = fill(P, cfg, rows, gene, replay, tree);
    if (rc) return rc;
    RL_ARG_CHECK(batch == 64 && sample_idx && ev_weight);
    P.batch = batch; P.t = t; P.sample_idx = sample_idx; P.ev_weight = ev_weight;
    k_sumtree_sample<<<cfg->n_worlds, 64, 0, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_sumtree_update(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                      const rl_sumtree_bufs* tree, int32_t batch, const int32_t* sample_idx, const float* errors,
                      void* stream) {
    TreeParams P;
    int rc = fill(P, cfg, rows, gene, replay, tree);
    if (rc) return rc;
    RL_ARG_CHECK(batch > 0 && sample_idx && errors);
    P.batch = batch; P.sample_idx = const_cast<int32_t*>(sample_idx); P.errors = errors;
    k_sumtree_update<<<(cfg->n_worlds * 32 + ST - 1) / ST, ST, 0, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_perdqn_epsilon_step(const rl_learn_bufs* learn, double* eps_dev, double eps_min, double eps_decay, void* stream) {
    RL_ARG_CHECK(learn && learn->grad && eps_dev && learn->kind == RL_MODEL_DQN);
    rl_model_dims d;
    int rc = rl_model_get_dims(learn->kind, &d);
    if (rc) return rc;
    k_perdqn_eps<<<1, 1, 0, (cudaStream_t)stream>>>(learn->grad + d.n_train, eps_dev, eps_min, eps_decay);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

}  // extern "C"

// replay_kernels.cu -- device replay rings, one per (world, brain).
//
// Reference: PrioritizedReplayBuffer (Models/PERD3QN.py:133-182): ring of `capacity` transitions, float32
// priorities initialised to 0, new items stored with max(priorities) (1.0 while empty), proportional sampling
// prio^0.6 with replacement (np.random.choice = cumsum + searchsorted), update_priorities by sequential
// overwrite.  The unused importance weights / beta bookkeeping (:168-172) are not materialised.
#include <cuda_fp16.h>
#include "rl_common.cuh"

namespace {

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {      // one F2FP.PACK_AB (full rate) instead of two F2F (quarter rate, XU pipe)
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));     // upper half <- b, lower half <- a
    return r;
}


constexpr int RT = 256;

struct ReplayParams {
    rl_world_cfg cfg;
    rl_world_bufs wb;
    rl_rows_bufs rows;
    rl_replay_bufs rp;
    int32_t gene, batch;
    uint64_t t;
    int32_t* sample_idx;
    const float* new_prio;
    int32_t* status;
    int32_t iter, n_iter, min_len;      // uniform sampler: draw stream of (event, iter); rings with len <= min_len are skipped
};

__device__ __forceinline__ float pw_of(float p) { return (float)pow((double)p, 0.6); }   // PERD3QN.py:162 (alpha)

__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane_id() == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < RT / 32; ++i) r = fmaxf(r, red[i]);
    __syncthreads();
    return r;
}

__device__ __forceinline__ int block_sum_i(int v, int* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane_id() == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = 0;
    for (int i = 0; i < RT / 32; ++i) r += red[i];
    __syncthreads();
    return r;
}

// ---- memorize: CTA per world ----
__global__ void __launch_bounds__(RT) k_replay_store(const ReplayParams P) {
    __shared__ float red[RT / 32];
    const int w = blockIdx.x, NW = P.cfg.n_worlds, S = P.cfg.slot_cap, ld = P.cfg.obs_ld, cap = P.rp.capacity;
    const int gk = P.gene * RL_N_ROW_KINDS + RL_ROWS_STORE;
    const int cnt = P.rows.count[(size_t)gk * NW + w];
    if (cnt == 0) return;
    const int off = P.rows.offset[(size_t)gk * NW + w];
    const int len = P.rp.len[w], pos = P.rp.pos[w];
    // max(priorities) (1.0 while the ring is empty, PERD3QN.py:147) and the number of entries holding it: read from the maintained
    // state when it is known, else one scan of the ring (first store after the priorities were written by hand / a count that fell to 0)
    __shared__ int s_more;
    float maxp = 1.0f;
    int mcount = 0;
    if (threadIdx.x == 0) s_more = 0;
    if (P.rp.prioritized && len > 0) {
        const int2 st = P.rp.maxst ? reinterpret_cast<const int2*>(P.rp.maxst)[w] : make_int2(0, 0);
        if (st.y > 0) { maxp = __int_as_float(st.x); mcount = st.y; }
        else {
            float m = 0.f;
            const float* pr = P.rp.prio + (size_t)w * cap;
            for (int i = threadIdx.x; i < cap; i += RT) m = fmaxf(m, pr[i]);
            maxp = block_max(m, red);
            int c = 0;
            for (int i = threadIdx.x; i < cap; i += RT) c += pr[i] == maxp;
            mcount = block_sum_i(c, reinterpret_cast<int*>(red));
        }
    }
    __syncthreads();
    const float pwv = P.rp.prioritized ? pw_of(maxp) : 1.0f;
    const int skip = max(0, cnt - cap);
    const int n_tr = min(cnt, P.rows.row_cap - off) - skip;        // transitions this launch writes
    // Phase 1: one thread per transition resolves the chain row id -> agent record (previous slot) and writes the scalars; the two
    // source rows and the ring slot go to shared memory.  Phase 2: the whole CTA copies the 2 x 40 float4 of every transition as one
    // flat loop of independent loads (the warp-per-transition chain of dependent loads left the kernel latency-bound at 0.15 ms).
    extern __shared__ int tr_src[];                                  // [3][slot_cap]: obs_state row, obs_prime row, ring slot
    int* src0 = tr_src; int* src1 = tr_src + S; int* dstq = tr_src + 2 * S;
    for (int k = threadIdx.x; k < n_tr; k += RT) {
        const int tr = skip + k;
        const int row = P.rows.rows[(size_t)gk * P.rows.row_cap + off + tr];
        const int4 rv = reinterpret_cast<const int4*>(P.wb.rec)[row];
        const int p = (pos + tr) % cap;
        src0[k] = w * S + ((rv.w >> 16) & 0xFFFF); src1[k] = row; dstq[k] = p;
        const size_t q = (size_t)w * cap + p;
        P.rp.action[q] = (int8_t)((rv.w >> 8) & 0xFF);
        P.rp.reward[q] = P.wb.reward[row];
        P.rp.done[q] = (rv.w & RL_F_DEAD) ? 1 : 0;
        if (P.rp.prioritized) {
            if (P.rp.prio[q] != maxp) atomicAdd(&s_more, 1);               // one more entry holds the maximum
            P.rp.prio[q] = maxp; P.rp.pw[q] = pwv;
        }
    }
    __syncthreads();
    const int nu = ld / 8;                                           // 32-byte units per row (ld = 160: 20)
    if (P.rp.obs_fp16 && P.wb.obs_state_h && P.wb.obs_prime_h) {
        // the World kernels wrote float16 copies of the rows (column 159 = 1.0 already): plain 16-byte copies, half the read traffic
        // (four independent loads in flight per thread before the first store)
        const int n_u = n_tr * 2 * nu;
        for (int i0 = threadIdx.x; i0 < n_u; i0 += 4 * RT) {
            uint4 v[4]; uint4* d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * RT;
                if (i < n_u) {
                    const int k = i / (2 * nu), rem = i - k * 2 * nu;
                    const int which = rem >= nu, u = rem - which * nu;
                    v[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(which ? P.wb.obs_prime_h : P.wb.obs_state_h) +
                                                                (size_t)(which ? src1[k] : src0[k]) * ld) + u);
                    d[j] = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(which ? P.rp.next_obs : P.rp.obs) + ((size_t)w * cap + dstq[k]) * ld) + u;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i0 + j * RT < n_u) *d[j] = v[j];
        }
    } else
    for (int i = threadIdx.x; i < n_tr * 2 * nu; i += RT) {
        const int k = i / (2 * nu), rem = i - k * 2 * nu;
        const int which = rem >= nu, u = rem - which * nu;
        const float4* sp = reinterpret_cast<const float4*>((which ? P.wb.obs_prime : P.wb.obs_state) + (size_t)(which ? src1[k] : src0[k]) * ld) + 2 * u;
        const float4 x = __ldg(sp);
        float4 y = __ldg(sp + 1);
        const size_t q = (size_t)w * cap + dstq[k];
        if (P.rp.obs_fp16) {        // float16 ring: rows rounded once here; the last (padding) column carries 1.0 (db1 rides the dW1 GEMM)
            if (u == nu - 1) y.w = 1.0f;
            uint4* h = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(which ? P.rp.next_obs : P.rp.obs) + q * ld);
            h[u] = make_uint4(pack_half2(x.x, x.y), pack_half2(x.z, x.w), pack_half2(y.x, y.y), pack_half2(y.z, y.w));
        } else {
            float4* d = reinterpret_cast<float4*>((which ? P.rp.next_obs : P.rp.obs) + q * ld) + 2 * u;
            d[0] = x; d[1] = y;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        P.rp.pos[w] = (pos + cnt) % cap;
        P.rp.len[w] = min(cap, len + cnt);
        if (P.rp.prioritized && P.rp.maxst) reinterpret_cast<int2*>(P.rp.maxst)[w] = make_int2(__float_as_int(maxp), mcount + s_more);
    }
}

// ---- sample: CTA per world, exact-integer CDF in shared memory ----
// floor(w * 2^24) of a non-negative float32 weight: the product is a pure exponent shift (exact in float32, weights are far below 2^39), so the
// single-precision conversion returns exactly what the float64 formulation of the oracle (int(float64(w) * 2^24)) does, without the FP64 pipe
__device__ __forceinline__ unsigned long long fix_of(float w) { return __float2ull_rz(w * 16777216.0f); }

__global__ void __launch_bounds__(RT) k_replay_sample(const ReplayParams P) {
    extern __shared__ __align__(16) unsigned long long cum[];
    __shared__ unsigned long long wsum[RT / 32];
    const int w = blockIdx.x, NW = P.cfg.n_worlds, cap = P.rp.capacity, batch = P.batch;
    const int gk = P.gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    const int cnt = P.rows.count[(size_t)gk * NW + w];
    if (cnt == 0) return;
    const int off = P.rows.offset[(size_t)gk * NW + w];
    const int len = P.rp.len[w];
    const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));
    unsigned long long total = 0;
    constexpr int ST = 10;                                  // tiles of 4 * RT weights held in registers (capacity <= 10240: the default 10000)
    if (P.rp.prioritized && (cap & 3) == 0 && len <= 4 * RT * ST) {
        // coalesced form: every thread loads its float4 of all tiles first (independent 16-byte loads), then the tiles are scanned
        // from registers -- thread-local prefix of 4, warp scan, running carry over the tiles.  Same integer sums, same cum[].
        const float4* pw4 = reinterpret_cast<const float4*>(P.rp.pw + (size_t)w * cap);
        const int lane = lane_id(), warp = threadIdx.x >> 5;
        float4 v[ST];
#pragma unroll
        for (int t = 0; t < ST; ++t) {
            const int i = (t * RT + (int)threadIdx.x) * 4;
            v[t] = i < len ? __ldg(pw4 + (i >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        unsigned long long carry = 0;
#pragma unroll
        for (int t = 0; t < ST; ++t) {
            if (t * RT * 4 >= len) break;                   // uniform
            const int i = (t * RT + (int)threadIdx.x) * 4;
            const unsigned long long f0 = i < len ? fix_of(v[t].x) : 0ull, f1 = i + 1 < len ? fix_of(v[t].y) : 0ull;
            const unsigned long long f2 = i + 2 < len ? fix_of(v[t].z) : 0ull, f3 = i + 3 < len ? fix_of(v[t].w) : 0ull;
            const unsigned long long local = f0 + f1 + f2 + f3;
            unsigned long long incl = local;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long x = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += x;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            unsigned long long base = carry, tile = 0;
#pragma unroll
            for (int k = 0; k < RT / 32; ++k) { const unsigned long long x = wsum[k]; tile += x; if (k < warp) base += x; }
            unsigned long long run = base + incl - local;
            if (i + 3 < len) {
                const unsigned long long c0 = run + f0, c1 = c0 + f1, c2 = c1 + f2;
                reinterpret_cast<ulonglong2*>(cum + i)[0] = make_ulonglong2(c0, c1);
                reinterpret_cast<ulonglong2*>(cum + i)[1] = make_ulonglong2(c2, c2 + f3);
            } else {
                if (i < len) { run += f0; cum[i] = run; }
                if (i + 1 < len) { run += f1; cum[i + 1] = run; }
                if (i + 2 < len) { run += f2; cum[i + 2] = run; }
            }
            carry += tile;
            __syncthreads();
        }
        total = carry;
    } else if (P.rp.prioritized) {
        const float* pw = P.rp.pw + (size_t)w * cap;
        const int per = (len + RT - 1) / RT;
        const int i0 = min(len, (int)threadIdx.x * per), i1 = min(len, i0 + per);
        unsigned long long local = 0;
        for (int i = i0; i < i1; ++i) local += fix_of(pw[i]);
        unsigned long long incl = local;
        const int lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned long long base = 0;
        for (int i = 0; i < warp; ++i) base += wsum[i];
        for (int i = 0; i < RT / 32; ++i) total += wsum[i];
        unsigned long long run = base + incl - local;
        for (int i = i0; i < i1; ++i) { run += fix_of(pw[i]); cum[i] = run; }
        __syncthreads();
    }
    for (int q = threadIdx.x; q < cnt * batch; q += RT) {
        const int e = q / batch;
        if (off + e >= P.rows.row_cap) break;
        const uint64_t bits = rl_draw(key, P.t, RL_SITE_REPLAY_SAMPLE, (uint32_t)q);
        int idx = 0;
        if (P.rp.prioritized) {
            if (total > 0) {
                const unsigned long long u53 = bits >> 11;
                const unsigned long long hi = __umul64hi(u53, total), lo = u53 * total;
                const unsigned long long target = (hi << 11) | (lo >> 53);     // floor(u53 * total / 2^53)
                int a = 0, b = len - 1;
                while (a < b) {
                    const int mid = (a + b) >> 1;
                    if (cum[mid] > target) b = mid; else a = mid + 1;
                }
                idx = a;
            }
        } else {
            idx = (int)rl_below(bits, (uint32_t)max(len, 1));
        }
        P.sample_idx[(size_t)(off + e) * batch + (q - e * batch)] = idx;
    }
}

// ---- uniform sampling WITHOUT replacement: random.sample(deque, k) (Models/D3QN.py:140, Models/DQN.py:100) ----
// CPython's algorithm (Lib/random.py, Random.sample) restated on the counter RNG: population index j counts from the
// OLDEST item of the deque (ring slot (pos - len + j) mod capacity).
//   n <= 277 (k in 6..85: setsize = 21 + 4^4): pool method -- j = below(n-i); result[i] = pool[j]; pool[j] = pool[n-i-1]
//   n  > 277: set method -- j = below(n), redrawn while j was already selected.
// Draw c of (event e, iteration it) is rl_draw(key, t, RL_SITE_REPLAY_SAMPLE_UNIFORM, ((e*n_iter + it) << 9) + c).
// One warp per event; the set method consumes 32 draws per round and accepts first occurrences in draw order, which
// is exactly the sequential accept/reject sequence because every draw is a pure function of its counter.
constexpr int POOL_MAX = 277;

__global__ void __launch_bounds__(RT) k_replay_sample_uniform(const ReplayParams P) {
    __shared__ uint16_t pool_s[RT / 32][POOL_MAX + 3];
    __shared__ int sel_s[RT / 32][128];
    const int w = blockIdx.x, NW = P.cfg.n_worlds, cap = P.rp.capacity, k = P.batch;
    const int gk = P.gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    const int cnt = P.rows.count[(size_t)gk * NW + w];
    if (cnt == 0) return;
    const int off = P.rows.offset[(size_t)gk * NW + w];
    const int n = P.rp.len[w], pos = P.rp.pos[w];
    const int first = ((pos - n) % cap + cap) % cap;                // ring slot of the oldest item
    const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    uint16_t* pool = pool_s[warp];
    int* sel = sel_s[warp];
    for (int e = warp; e < cnt; e += RT / 32) {
        if (off + e >= P.rows.row_cap) break;
        int32_t* out = P.sample_idx + (size_t)(off + e) * k;
        const uint32_t base = ((uint32_t)(e * P.n_iter + P.iter)) << 9;
        if (n <= P.min_len || n < k) {                              // skipped event (DQN.py:79) / reference raises ValueError
            for (int i = lane; i < k; i += 32) out[i] = -1;
            if (n > P.min_len && lane == 0 && P.status) atomicOr(P.status, 1);   // random.sample would raise ValueError
            continue;
        }
        if (n <= POOL_MAX) {
            for (int i = lane; i < n; i += 32) pool[i] = (uint16_t)i;
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < k; ++i) {
                    const int j = (int)rl_below(rl_draw(key, P.t, RL_SITE_REPLAY_SAMPLE_UNIFORM, base + i), (uint32_t)(n - i));
                    sel[i] = pool[j];
                    pool[j] = pool[n - i - 1];
                }
            }
            __syncwarp();
        } else {
            int nsel = 0;
            for (uint32_t c = 0; nsel < k; c += 32) {
                const int j = (int)rl_below(rl_draw(key, P.t, RL_SITE_REPLAY_SAMPLE_UNIFORM, base + c + lane), (uint32_t)n);
                bool dup = false;
                for (int i = 0; i < nsel; ++i) dup |= sel[i] == j;
                const unsigned m = __match_any_sync(0xffffffffu, j);
                const bool ok = !dup && (__ffs(m) - 1) == lane;
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                const int rank = nsel + __popc(bal & lanemask_lt());
                if (ok && rank < k) sel[rank] = j;
                nsel = min(k, nsel + __popc(bal));
                __syncwarp();
            }
        }
        for (int i = lane; i < k; i += 32) { int p = first + sel[i]; out[i] = p >= cap ? p - cap : p; }
        __syncwarp();
    }
}

// ---- update_priorities: warp per world, sequential over events, later writes win ----
// Also keeps {max(priorities), count of entries holding it} exact (rl_replay_bufs.maxst): a per-warp bitmap in shared memory marks
// the distinct touched entries; before the writes those holding the maximum are counted, after the writes the final values give the
// new maximum / count.  Rings too large for the bitmap (or whose state is unknown) are left "unknown": the next store scans.
constexpr int UP_BITMAP_WORDS = 1024;        // per warp: capacity <= 32768
constexpr int UP_K = 10;                     // entries per lane and chunk: 320 = five 64-row events, all their loads in flight at once
__global__ void __launch_bounds__(RT) k_replay_update_prio(const ReplayParams P) {
    __shared__ uint32_t bitmap[RT / 32][UP_BITMAP_WORDS];
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= P.cfg.n_worlds) return;
    const int NW = P.cfg.n_worlds, cap = P.rp.capacity, batch = P.batch, lane = lane_id();
    const int gk = P.gene * RL_N_ROW_KINDS + RL_ROWS_EVENT;
    int cnt = P.rows.count[(size_t)gk * NW + w];
    if (cnt == 0) return;
    const int off = P.rows.offset[(size_t)gk * NW + w];
    cnt = min(cnt, P.rows.row_cap - off);
    const int n_all = cnt * batch;                       // entries of this world, in event order (later writes win)
    const int32_t* sidx = P.sample_idx + (size_t)off * batch;
    const float* nprio = P.new_prio + (size_t)off * batch;
    float* prio = P.rp.prio + (size_t)w * cap;
    float* pw = P.rp.pw + (size_t)w * cap;
    uint32_t* bm = bitmap[threadIdx.x >> 5];
    const int2 st = P.rp.maxst ? reinterpret_cast<const int2*>(P.rp.maxst)[w] : make_int2(0, 0);
    const bool track = P.rp.maxst && st.y > 0 && cap <= 32 * UP_BITMAP_WORDS;
    const float M = __int_as_float(st.x);
    int dec = 0;
    if (track) {     // pass A: the distinct entries about to be overwritten that hold the maximum (old values, before any write)
        for (int k = lane; k < (cap + 31) / 32; k += 32) bm[k] = 0u;
        __syncwarp();
        for (int c0 = 0; c0 < n_all; c0 += 32 * UP_K) {
            int idx[UP_K]; float old[UP_K];
#pragma unroll
            for (int k = 0; k < UP_K; ++k) { const int j = c0 + k * 32 + lane; idx[k] = j < n_all ? sidx[j] : -1; }
#pragma unroll
            for (int k = 0; k < UP_K; ++k) old[k] = idx[k] >= 0 ? prio[idx[k]] : 0.f;
#pragma unroll
            for (int k = 0; k < UP_K; ++k)
                if (idx[k] >= 0) {
                    const uint32_t bit = 1u << (idx[k] & 31);
                    if (!(atomicOr(&bm[idx[k] >> 5], bit) & bit) && old[k] == M) ++dec;      // first mention of the entry
                }
        }
        __syncwarp();
    }
    for (int c0 = 0; c0 < n_all; c0 += 32 * UP_K) {      // the writes, 32 entries at a time in order; inside a group the last lane wins
        int idx[UP_K]; float val[UP_K];
#pragma unroll
        for (int k = 0; k < UP_K; ++k) {
            const int j = c0 + k * 32 + lane;
            idx[k] = j < n_all ? sidx[j] : -1;
            val[k] = j < n_all ? nprio[j] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < UP_K; ++k) {
            if (c0 + k * 32 >= n_all) break;                                                  // (warp-uniform)
            const unsigned m = __match_any_sync(0xffffffffu, idx[k] >= 0 ? idx[k] : -1 - lane);
            if (idx[k] >= 0 && (31 - __clz(m)) == lane) { prio[idx[k]] = val[k]; pw[idx[k]] = pw_of(val[k]); }
            __syncwarp();
        }
    }
    if (!P.rp.maxst) return;
    if (!track) {                                          // unknown stays unknown; a known state this kernel could not follow is dropped
        if (lane == 0 && st.y > 0) reinterpret_cast<int2*>(P.rp.maxst)[w] = make_int2(0, 0);
        return;
    }
    __threadfence_block();
    __syncwarp();
    float vmax = -1.0f; int nmax = 0, nM = 0;              // pass C, over the distinct touched entries: largest final value, how many hold it / hold M
    for (int c0 = 0; c0 < n_all; c0 += 32 * UP_K) {
        int idx[UP_K]; float fin[UP_K];
#pragma unroll
        for (int k = 0; k < UP_K; ++k) { const int j = c0 + k * 32 + lane; idx[k] = j < n_all ? sidx[j] : -1; }
#pragma unroll
        for (int k = 0; k < UP_K; ++k) fin[k] = idx[k] >= 0 ? prio[idx[k]] : 0.f;
#pragma unroll
        for (int k = 0; k < UP_K; ++k)
            if (idx[k] >= 0) {
                const uint32_t bit = 1u << (idx[k] & 31);
                if (atomicAnd(&bm[idx[k] >> 5], ~bit) & bit) {    // first to clear the mark: one visit per distinct entry
                    const float v = fin[k];
                    if (v > vmax) { vmax = v; nmax = 1; } else if (v == vmax) ++nmax;
                    nM += v == M;
                }
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, vmax, o);
        const int on = __shfl_xor_sync(0xffffffffu, nmax, o);
        if (ov > vmax) { vmax = ov; nmax = on; } else if (ov == vmax) nmax += on;
        nM += __shfl_xor_sync(0xffffffffu, nM, o);
        dec += __shfl_xor_sync(0xffffffffu, dec, o);
    }
    if (lane == 0) {
        int2 out;
        if (vmax > M) out = make_int2(__float_as_int(vmax), nmax);
        else out = make_int2(st.x, max(0, st.y - dec + nM));
        reinterpret_cast<int2*>(P.rp.maxst)[w] = out;
    }
}

int fill(ReplayParams& P, const rl_world_cfg* cfg, const rl_world_bufs* wb, const rl_rows_bufs* rows, int32_t gene,
         const rl_replay_bufs* rp) {
    RL_ARG_CHECK(cfg && rows && rp);
    RL_ARG_CHECK(gene >= 0 && gene < cfg->n_genes);
    RL_ARG_CHECK(rp->capacity > 0 && rp->len && rp->pos);
    P.cfg = *cfg;
    if (wb) P.wb = *wb;
    P.rows = *rows; P.rp = *rp; P.gene = gene; P.batch = 0; P.t = 0; P.sample_idx = nullptr; P.new_prio = nullptr; P.iter = 0; P.n_iter = 1; P.min_len = 0; P.status = nullptr;
    return RL_OK;
}

}  // namespace

extern "C" {

int rl_replay_store(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                    const rl_replay_bufs* replay, void* stream) {
    ReplayParams P;
    RL_ARG_CHECK(bufs);
    int rc = fill(P, cfg, bufs, rows, gene, replay);
    if (rc) return rc;
    RL_ARG_CHECK(replay->obs && replay->next_obs && replay->action && replay->reward && replay->done);
    RL_ARG_CHECK(!replay->prioritized || (replay->prio && replay->pw));
    k_replay_store<<<cfg->n_worlds, RT, 3 * sizeof(int) * (size_t)cfg->slot_cap, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_replay_sample(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                     int32_t batch, uint64_t t, int32_t* sample_idx, void* stream) {
    ReplayParams P;
    int rc = fill(P, cfg, nullptr, rows, gene, replay);
    if (rc) return rc;
    RL_ARG_CHECK(batch > 0 && sample_idx);
    P.batch = batch; P.t = t; P.sample_idx = sample_idx;
    const size_t smem = replay->prioritized ? (size_t)replay->capacity * 8 : 0;
    if (smem > 200 * 1024) return rl_set_err(RL_ERR_UNSUPPORTED, "prioritized capacity %d exceeds the in-SM CDF (25600)", replay->capacity);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_replay_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    k_replay_sample<<<cfg->n_worlds, RT, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_replay_sample_uniform(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                             int32_t batch, uint64_t t, int32_t iter, int32_t n_iter, int32_t min_len, int32_t* sample_idx,
                             int32_t* status, void* stream) {
    ReplayParams P;
    int rc = fill(P, cfg, nullptr, rows, gene, replay);
    if (rc) return rc;
    RL_ARG_CHECK(batch > 5 && batch <= 85 && sample_idx);          // setsize = 277 holds for 6 <= k <= 85 (Lib/random.py)
    RL_ARG_CHECK(n_iter > 0 && iter >= 0 && iter < n_iter && min_len >= 0);
    P.batch = batch; P.t = t; P.sample_idx = sample_idx; P.iter = iter; P.n_iter = n_iter; P.min_len = min_len; P.status = status;
    k_replay_sample_uniform<<<cfg->n_worlds, RT, 0, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_replay_update_prio(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                          int32_t batch, const int32_t* sample_idx, const float* new_prio, void* stream) {
    ReplayParams P;
    int rc = fill(P, cfg, nullptr, rows, gene, replay);
    if (rc) return rc;
    RL_ARG_CHECK(batch > 0 && sample_idx && new_prio);
    if (!replay->prioritized) return RL_OK;
    P.batch = batch; P.sample_idx = const_cast<int32_t*>(sample_idx); P.new_prio = new_prio;
    k_replay_update_prio<<<(cfg->n_worlds * 32 + RT - 1) / RT, RT, 0, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

}  // extern "C"

// world_kernels.cu -- ReinLife World hot path on sm_100a: one CTA per world, the whole world
// (cell types + agent attributes) staged in shared memory, all order-dependent reference loops
// replaced by closed-form gather rules (no atomics on the data path), sequential RNG placements
// done by warp 0 on ballot-built bitmasks, observations written as full 16-byte-vector rows.
//
// Reference semantics: ReinLife/World/environment.py (cited per phase), grid.py:60-117,
// entities.py:145-248; closed forms: SURVEY.md Appendix A.3-A.6 (validated against the reference
// through oracle/rl_oracle.c, which keeps the sequential formulation).
#include <stdlib.h>
#include <type_traits>
#include <cuda_fp16.h>
#include "rl_common.cuh"

namespace {

#ifndef RL_WT
#define RL_WT 256
#endif
constexpr int WT = RL_WT;       // threads per world
#ifndef RL_WORLD_MINB
#define RL_WORLD_MINB 6
#endif
constexpr int WNW = WT / 32;    // warps per world


enum { M_ALIVE = 0, M_NFOOD, M_NPOISON, M_NSUPER, M_NB, M_PRESENT, M_ANY, M_PALL, M_ALIVE_G = 8,
       M_CNT_G = M_ALIVE_G + RL_MAX_GENES, M_PG = M_CNT_G + RL_MAX_GENES, M_WORDS = M_PG + RL_MAX_GENES };

// non-static families (static_families = False; environment.py:149,506-507,541-547,728-739): genes are unbounded lineage
// numbers, so inside a kernel every distinct gene of the world gets a compact id < NS_IDS (shared-memory hash table, ids by
// table position: deterministic); per-gene tables are indexed by that id and rec[].gene carries the real gene.
constexpr int NS_IDS = 256;     // distinct lineages alive in one world at a time (status bit 2 when exceeded)
constexpr int NS_HASH = 1024;

struct WParams {
    long long* trace;
    rl_world_cfg cfg;
    rl_world_bufs b;
    rl_world_ns_bufs ns;
    uint64_t t;
    int32_t target, max_age, which;
    int32_t dbg;        // RL_WORLD_DEBUG bits (timing experiments only): 1 skip row stores, 2 skip observe, 4 early exit after sim
    uint32_t magicW;
};

struct WS {
    uint32_t* planes;   // [2][(H+6)*(W+6)] toroidally padded observation planes: [0] half2 {food, dead-agent gene or -2}
                        // (all values exact in fp16), [1] float health ratio
    float* ascal;       // [WNW][16][8] per-warp batch of agent scalars (6 observation scalars + 2 zeros)
    uint32_t* mask;     // [Cw] empty cells
    uint32_t* amask;    // [Cw] agent cells / eligible parents
    uint32_t* wpre;     // [Cw] exclusive popc prefix of amask
    int32_t* misc;      // [M_WORDS]
    int32_t* offtab;    // [160] row element e -> offset into the padded planes (window elements), 0 for scalars/pad
    int32_t* rowmap;    // [H+6] padded row -> source row * W     (Grid.fov's toroidal concatenate, grid.py:99-115)
    int32_t* colmap;    // [W+6] padded column -> source column
    float* rtab;        // [RL_MAX_GENES][8] reward by (gene, dead, killed) (_get_rewards, environment.py:291-311); [4..7] = reward / 100.0 (PPO.py:73)
    int16_t *health, *age, *maxage;
    uint16_t *aslot, *tgt, *src, *cellof;
    uint8_t *type, *ntype, *flags, *gene;
    int8_t* action;
    int32_t *alive_g, *cnt_g, *pg;   // per-gene (static) / per-lineage-id (non-static) tables: alive count, listed count, count / n (float bits)
    // ---- non-static families only ----
    double* rtab64;     // [NS_IDS][4] float64 reward by (id, dead, killed): Agent.fitness accumulates it (entities.py:187-192)
    double* fit;        // [slot_cap] Agent.fitness by OLD slot
    long long* ser;     // [slot_cap] object identity by OLD slot
    int32_t* hkey;      // [NS_HASH] gene keys of the id hash table (-1 = empty)
    int32_t* id2gene;   // [NS_IDS]
    uint8_t* hid;       // [NS_HASH] id of a table entry
    uint32_t* nbmask;   // [Cw] cells holding a newborn agent (serial not handed out yet)
    uint32_t* nbpre;    // [Cw] exclusive popc prefix of nbmask
    rl_ns_state* st;    // the world's max_gene / best_agents / last _produce event
};

__host__ __device__ inline size_t ws_bytes_ns(int H, int W, int S) {
    const size_t C = (size_t)H * W, Cw = ((C + 31) / 32 + 3) & ~(size_t)3;
    const size_t Sp = ((size_t)S + 1) & ~(size_t)1;
    return 4 * (size_t)NS_IDS * 8 + 8 * (size_t)NS_IDS * 4 + 16 * Sp + 4 * NS_HASH + 4 * NS_IDS + NS_HASH + 4 * 3 * NS_IDS + 4 * Cw * 2 + sizeof(rl_ns_state) + 64;
}

__host__ __device__ inline size_t ws_bytes(int H, int W) {
    const size_t C = (size_t)H * W, Cw = (C + 31) / 32, Cp = (C + 15) & ~(size_t)15;
    const size_t PADN = ((size_t)(H + 6) * (W + 6) + 3) & ~(size_t)3;
    const size_t TAB = 160 + (((size_t)H + 6 + 3) & ~(size_t)3) + (((size_t)W + 6 + 3) & ~(size_t)3) + RL_MAX_GENES * 8;
    return 4 * 2 * PADN + 4 * WNW * 16 * 8 + 4 * ((Cw + 3) & ~(size_t)3) * 3 + 4 * M_WORDS + 4 * TAB + 2 * Cp * 7 + Cp * 5;
}

// the non-static additions sit behind the static carve-up (8-byte aligned)
__device__ inline void ws_carve_ns(WS& s, unsigned char* base, int H, int W, int S) {
    const size_t C = (size_t)H * W, Cw = ((C + 31) / 32 + 3) & ~(size_t)3;
    const size_t Sp = ((size_t)S + 1) & ~(size_t)1;
    unsigned char* p = base + ((ws_bytes(H, W) + 15) & ~(size_t)15);
    s.rtab64 = (double*)p; p += 8 * (size_t)NS_IDS * 4;
    s.fit = (double*)p; p += 8 * Sp;
    s.ser = (long long*)p; p += 8 * Sp;
    s.st = (rl_ns_state*)p; p += (sizeof(rl_ns_state) + 15) & ~(size_t)15;
    s.rtab = (float*)p; p += 4 * (size_t)NS_IDS * 8;
    s.hkey = (int32_t*)p; p += 4 * NS_HASH;
    s.id2gene = (int32_t*)p; p += 4 * NS_IDS;
    s.alive_g = (int32_t*)p; p += 4 * NS_IDS;
    s.cnt_g = (int32_t*)p; p += 4 * NS_IDS;
    s.pg = (int32_t*)p; p += 4 * NS_IDS;
    s.nbmask = (uint32_t*)p; p += 4 * Cw;
    s.nbpre = (uint32_t*)p; p += 4 * Cw;
    s.hid = p;
}

__device__ inline void ws_carve(WS& s, unsigned char* base, int H, int W) {
    const size_t C = (size_t)H * W;
    const size_t Cw = ((C + 31) / 32 + 3) & ~(size_t)3;
    const size_t Cp = (C + 15) & ~(size_t)15;
    const size_t PADN = ((size_t)(H + 6) * (W + 6) + 3) & ~(size_t)3;
    unsigned char* p = base;
    s.planes = (uint32_t*)p; p += 4 * 2 * PADN;
    s.ascal = (float*)p; p += 4 * WNW * 16 * 8;
    s.mask = (uint32_t*)p; p += 4 * Cw;
    s.amask = (uint32_t*)p; p += 4 * Cw;
    s.wpre = (uint32_t*)p; p += 4 * Cw;
    s.misc = (int32_t*)p; p += 4 * M_WORDS;
    s.offtab = (int32_t*)p; p += 4 * 160;
    s.rowmap = (int32_t*)p; p += 4 * (((size_t)H + 6 + 3) & ~(size_t)3);
    s.colmap = (int32_t*)p; p += 4 * (((size_t)W + 6 + 3) & ~(size_t)3);
    s.rtab = (float*)p; p += 4 * RL_MAX_GENES * 8;
    s.health = (int16_t*)p; p += 2 * Cp;
    s.age = (int16_t*)p; p += 2 * Cp;
    s.maxage = (int16_t*)p; p += 2 * Cp;
    s.aslot = (uint16_t*)p; p += 2 * Cp;
    s.tgt = (uint16_t*)p; p += 2 * Cp;
    s.src = (uint16_t*)p; p += 2 * Cp;
    s.cellof = (uint16_t*)p; p += 2 * Cp;
    s.type = p; p += Cp;
    s.ntype = p; p += Cp;
    s.flags = p; p += Cp;
    s.gene = p; p += Cp;
    s.action = (int8_t*)p;
    s.alive_g = s.misc + M_ALIVE_G; s.cnt_g = s.misc + M_CNT_G; s.pg = s.misc + M_PG;
}

// toroidal neighbour: 0 up (i-1), 1 right (j+1), 2 down (i+1), 3 left (j-1) -- environment.py:601-623
__device__ __forceinline__ int nbr(int c, int d, int W, int C, uint32_t magicW) {
    if (d == 0) return c < W ? c + C - W : c - W;
    if (d == 2) return c + W >= C ? c + W - C : c + W;
    int i = (int)__umulhi((uint32_t)c, magicW);
    int j = c - i * W;
    if (d == 1) return j == W - 1 ? c - (W - 1) : c + 1;
    return j == 0 ? c + (W - 1) : c - 1;
}

// ---- warp-0 sequential placement on the empty-cell bitmask (Grid.set_random, grid.py:69-83) ----
// Returns the cell holding the k-th empty cell in row-major order, k = rl_below(bits, n_empty); -1 if none.
__device__ int warp_place(uint32_t* mask, int nwords, uint64_t bits) {
    const int lane = lane_id();
    const int wpl = (nwords + 31) / 32;
    const int w0 = lane * wpl, w1 = min(nwords, w0 + wpl);
    int cnt = 0;
    for (int wd = w0; wd < w1; ++wd) cnt += __popc(mask[wd]);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return -1;
    const int k = (int)rl_below(bits, (uint32_t)total);
    const unsigned bal = __ballot_sync(0xffffffffu, incl > k);
    const int owner = __ffs(bal) - 1;
    int cell = -1;
    if (lane == owner) {
        int kk = k - (incl - cnt);
        for (int wd = w0; wd < w1; ++wd) {
            int p = __popc(mask[wd]);
            if (kk < p) { cell = wd * 32 + (int)__fns(mask[wd], 0, kk + 1); break; }
            kk -= p;
        }
    }
    cell = __shfl_sync(0xffffffffu, cell, owner);
    return cell;
}

__device__ __forceinline__ void warp_mask_clear(uint32_t* mask, int cell) {
    if (lane_id() == 0) mask[cell >> 5] &= ~(1u << (cell & 31));
    __syncwarp();
}

// The same placement with the popc prefix kept in registers between calls (the hot kernels place several entities per phase):
// lane l owns the words [l * wpl, (l + 1) * wpl); a placement takes one empty cell out of its owner's count and out of the
// inclusive prefixes of the lanes behind it, and clears the cell's bit in `mask`.  Same k, same cell as warp_place.
struct Placer { int cnt, incl; };
__device__ __forceinline__ Placer placer_init(const uint32_t* mask, int nwords) {
    const int lane = lane_id();
    const int wpl = (nwords + 31) / 32;
    const int w0 = lane * wpl, w1 = min(nwords, w0 + wpl);
    Placer pl; pl.cnt = 0;
    for (int wd = w0; wd < w1; ++wd) pl.cnt += __popc(mask[wd]);
    pl.incl = pl.cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, pl.incl, o);
        if (lane >= o) pl.incl += v;
    }
    return pl;
}
__device__ __forceinline__ int placer_take(Placer& pl, uint32_t* mask, int nwords, uint64_t bits) {
    const int lane = lane_id();
    const int total = __shfl_sync(0xffffffffu, pl.incl, 31);
    if (total == 0) return -1;
    const int k = (int)rl_below(bits, (uint32_t)total);
    const int owner = __ffs(__ballot_sync(0xffffffffu, pl.incl > k)) - 1;
    int cell = -1;
    if (lane == owner) {
        const int wpl = (nwords + 31) / 32;
        int kk = k - (pl.incl - pl.cnt);
        for (int wd = lane * wpl; ; ++wd) {
            const uint32_t m = mask[wd];
            const int p = __popc(m);
            if (kk < p) {
                const int b = (int)__fns(m, 0, kk + 1);
                mask[wd] = m & ~(1u << b);
                cell = wd * 32 + b;
                break;
            }
            kk -= p;
        }
        --pl.cnt;
    }
    if (lane >= owner) --pl.incl;
    cell = __shfl_sync(0xffffffffu, cell, owner);
    return cell;
}

__device__ __forceinline__ void spawn_agent(WS& s, int cell, int gene, int health, int age) {   // entities.py:145-160
    if (lane_id() == 0) {
        s.type[cell] = RL_AGENT;
        s.health[cell] = (int16_t)health; s.age[cell] = (int16_t)age; s.maxage[cell] = 50;
        s.gene[cell] = (uint8_t)gene; s.flags[cell] = 0; s.action[cell] = -1; s.aslot[cell] = RL_NONE16;
    }
    __syncwarp();
}

// build the empty-cell bitmask of `tarr` (all warps)
__device__ __forceinline__ void build_empty_mask(const uint8_t* tarr, uint32_t* mask, int C) {
    const int Cw = (C + 31) / 32;
    for (int ch = threadIdx.x >> 5; ch < Cw; ch += WNW) {
        int c = ch * 32 + lane_id();
        unsigned m = __ballot_sync(0xffffffffu, c < C && tarr[c] == RL_EMPTY);
        if (lane_id() == 0) mask[ch] = m;
    }
}

// per-CTA lookup tables of the observation phase (pure functions of H, W): computed once by the first threads
__device__ __forceinline__ void init_tables(const WParams& P, WS& s) {
    const int H = P.cfg.height, W = P.cfg.width, PW = W + 6;
    const int PADN = (((H + 6) * PW) + 3) & ~3;
    const int t = threadIdx.x;
    for (int e = t; e < 160; e += WT) {
        const int pl = e / 49, q = e - pl * 49;
        s.offtab[e] = e < 147 ? (pl == 1 ? PADN : 0) + (q / 7) * PW + (q - (q / 7) * 7) : 0;   // food / gene share plane 0
    }
    for (int k = t; k < H + 6; k += WT) { int si = k - 3; si += si < 0 ? H : 0; si -= si >= H ? H : 0; s.rowmap[k] = si * W; }   // H, W >= 3
    for (int k = t; k < W + 6; k += WT) { int sj = k - 3; sj += sj < 0 ? W : 0; sj -= sj >= W ? W : 0; s.colmap[k] = sj; }
}

constexpr int M_NIDS = M_PRESENT;      // non-static: number of lineage ids in use (M_PRESENT is a static-families scratch word)
constexpr int M_NLIN = M_ALIVE_G;      // non-static: distinct lineages in the final list (the per-gene words of misc are unused)
constexpr int M_NNEW = M_ANY;          // non-static: newborn agents of this phase

__device__ __forceinline__ uint32_t ns_hash(int gene) { return ((uint32_t)gene * 2654435761u) >> 22; }      // 10 bits = NS_HASH
__device__ __forceinline__ int ns_lookup(const WS& s, int gene) {
    uint32_t h = ns_hash(gene);
    while (s.hkey[h] != gene) h = (h + 1) & (NS_HASH - 1);
    return s.hid[h];
}

// load cell types + the agent list into the cell-indexed shared arrays
template <bool DECAY, bool NS>
__device__ __forceinline__ int load_world(const WParams& P, WS& s, int w) {
    const int C = P.cfg.height * P.cfg.width;
    const uint8_t* tg = P.b.type + (size_t)w * C;
    if ((C & 3) == 0) {                                    // rows of the type array are 4-byte aligned
        const uint32_t* tg4 = reinterpret_cast<const uint32_t*>(tg);
        for (int q = threadIdx.x; q < C / 4; q += WT) {
            reinterpret_cast<uint32_t*>(s.type)[q] = __ldg(tg4 + q);
            reinterpret_cast<uint2*>(s.aslot)[q] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
        }
    } else {
        for (int c = threadIdx.x; c < C; c += WT) { s.type[c] = tg[c]; s.aslot[c] = RL_NONE16; }
    }
    for (int k = threadIdx.x; k < M_WORDS; k += WT) s.misc[k] = 0;
    init_tables(P, s);
    const int n = min(P.b.n_agents[w], P.cfg.slot_cap);
    const int4* rg = reinterpret_cast<const int4*>(P.b.rec + (size_t)w * P.cfg.slot_cap);
    if (NS) {
        for (int k = threadIdx.x; k < NS_HASH; k += WT) s.hkey[k] = -1;
        for (int k = threadIdx.x; k < NS_IDS; k += WT) { s.alive_g[k] = 0; s.cnt_g[k] = 0; s.id2gene[k] = -1; }
        for (int k = threadIdx.x; k < (int)(sizeof(rl_ns_state) / 8); k += WT)
            reinterpret_cast<long long*>(s.st)[k] = reinterpret_cast<const long long*>(P.ns.state + w)[k];
        __syncthreads();
        for (int sl = threadIdx.x; sl < n; sl += WT) {     // phase 1: the distinct genes of the list
            const int gene = rg[sl].z;
            uint32_t h = ns_hash(gene);
            for (;;) {
                const int old = atomicCAS(&s.hkey[h], -1, gene);
                if (old == -1 || old == gene) break;
                h = (h + 1) & (NS_HASH - 1);
            }
        }
        __syncthreads();
        {                                                  // phase 2: ids by table position (deterministic)
            constexpr int PER = NS_HASH / WT;
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < PER; ++k) cnt += s.hkey[threadIdx.x * PER + k] != -1;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane_id() >= o) incl += v; }
            int* wsum = reinterpret_cast<int*>(s.rtab64);                // (scratch: the reward table is built later)
            if (lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
            __syncthreads();
            int base = incl - cnt;
            for (int k = 0; k < (int)(threadIdx.x >> 5); ++k) base += wsum[k];
            for (int k = 0; k < PER; ++k) {
                const int e = threadIdx.x * PER + k;
                if (s.hkey[e] != -1) {
                    if (base < NS_IDS) { s.hid[e] = (uint8_t)base; s.id2gene[base] = s.hkey[e]; }
                    else { s.hid[e] = (uint8_t)(NS_IDS - 1); if (P.b.status) atomicOr(&P.b.status[w], 4); }
                    ++base;
                }
            }
            if (threadIdx.x == WT - 1) s.misc[M_NIDS] = min(base, NS_IDS);
        }
    }
    __syncthreads();
    for (int sl = threadIdx.x; sl < n; sl += WT) {
        int4 v = ld_stream_i4(rg + sl);
        int c = v.x & 0xFFFF;
        int h = (int16_t)(v.x >> 16), age = (int16_t)(v.y & 0xFFFF), ma = (int16_t)(v.y >> 16);
        int fl = v.w & 0xFF, act = (int8_t)((v.w >> 8) & 0xFF);
        if (DECAY) {                                       // _act, environment.py:268-271
            h = min(200, h - 10);
            age = min(ma, age + 1);
            fl &= ~(RL_F_KILLED | RL_F_INTER_KILLED | RL_F_INTRA_KILLED);
        }
        s.health[c] = (int16_t)h; s.age[c] = (int16_t)age; s.maxage[c] = (int16_t)ma;
        s.gene[c] = NS ? (uint8_t)ns_lookup(s, v.z) : (uint8_t)v.z; s.flags[c] = (uint8_t)fl; s.action[c] = (int8_t)act;
        s.aslot[c] = (uint16_t)sl; s.cellof[sl] = (uint16_t)c;
        if (NS) {
            s.fit[sl] = P.ns.fitness[(size_t)w * P.cfg.slot_cap + sl];
            s.ser[sl] = P.ns.serial[(size_t)w * P.cfg.slot_cap + sl];
        }
    }
    __syncthreads();
    return n;
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {      // one F2FP.PACK_AB
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// float32(health / max_health), environment.py:365,397.  The correctly rounded float32 quotient equals
// float32(float64(h) / 200.0) for every int16 h (no double-rounding case exists; checked exhaustively), so no f64 here.
__device__ __forceinline__ float hratio(int h) { return __fdiv_rn((float)h, 200.0f); }

// Rebuild the agent list from the final grid `ft` (Grid.get_entities, grid.py:60-67), write
// type/rec/n_agents(/reward), then Environment._get_observations (environment.py:313-375).
// src[c] = index into the attribute arrays of the agent standing on cell c.
//
// Observation: the three planes are materialised ONCE per world as floats on a toroidally padded (H+6)x(W+6)
// grid (what Grid.fov's np.concatenate builds, grid.py:99-115), so a window element is one shared-memory load at
// base(i,j) + constant(lane): each lane owns row elements e = lane + 32k (k < 5) of the 160-float row and
// the warp writes the row as five fully coalesced, 128-byte aligned stores straight to HBM.
template <bool STEP, bool NS>
__device__ void finish_and_observe(const WParams& P, WS& s, int w, const uint8_t* ft, float* obs_out, int n_prev) {
    const int H = P.cfg.height, W = P.cfg.width, C = H * W, Cw = (C + 31) / 32;
    const int S = P.cfg.slot_cap, ld = P.cfg.obs_ld, G = P.cfg.n_genes;
    const int PW = W + 6, PADN = (((H + 6) * PW) + 3) & ~3;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    uint8_t* tg = P.b.type + (size_t)w * C;

    for (int ch = warp; ch < Cw; ch += WNW) {
        int c = ch * 32 + lane;
        bool in = c < C;
        uint8_t t = in ? ft[c] : (uint8_t)0;
        unsigned m = __ballot_sync(0xffffffffu, in && t == RL_AGENT);
        if (lane == 0) s.amask[ch] = m;
        if (NS) {                                          // newborn = no slot in the previous list: its serial is handed out below
            const unsigned nb = __ballot_sync(0xffffffffu, in && t == RL_AGENT && s.aslot[s.src[c]] == RL_NONE16);
            if (lane == 0) s.nbmask[ch] = nb;
        }
        if ((C & 3) != 0 && in) tg[c] = t;
    }
    if ((C & 3) == 0)
        for (int q = threadIdx.x; q < C / 4; q += WT) reinterpret_cast<uint32_t*>(tg)[q] = reinterpret_cast<const uint32_t*>(ft)[q];
    const int n_tab = NS ? s.misc[M_NIDS] : P.cfg.n_genes;
    if (STEP) {                                          // _get_rewards (environment.py:291-311) by (gene, dead, killed)
        for (int g = threadIdx.x; g < n_tab; g += WT) {
            const int alive_ = s.misc[M_ALIVE], kin = max(0, s.alive_g[g] - 1);
            const double ra = alive_ == 1 ? 0.0 : (double)kin / (double)max(alive_, 1), rd = (double)(kin - alive_);
            const double bonus = P.cfg.incentivize_killing ? 0.2 : 0.0;
            const double r4[4] = {ra, ra + bonus, rd, rd + bonus};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s.rtab[g * 8 + k] = (float)r4[k]; s.rtab[g * 8 + 4 + k] = (float)(r4[k] / 100.0);
                if (NS) s.rtab64[g * 4 + k] = r4[k];
            }
        }
    }
    __syncthreads();
    if (STEP && NS) {
        // Agent.fitness += reward for every agent of the _act list, vanished ones included (entities.py:187-192); best_agents
        // holds live objects, so an entry that is one of these agents sees the new fitness (environment.py:149,728-739)
        for (int sl = threadIdx.x; sl < n_prev; sl += WT) {
            const int c0 = s.cellof[sl];
            const unsigned fl = s.flags[c0];
            const double f = s.fit[sl] + s.rtab64[s.gene[c0] * 4 + ((fl & RL_F_DEAD) ? 2 : 0) + ((fl & RL_F_KILLED) ? 1 : 0)];
            s.fit[sl] = f;
            const long long id = s.ser[sl];
#pragma unroll
            for (int k = 0; k < 10; ++k) if (s.st->best[k].serial == id) s.st->best[k].fitness = f;
        }
        __syncthreads();
    }
    if (warp == 0) {
        const int wpl = (Cw + 31) / 32;
        const int w0 = lane * wpl, w1 = min(Cw, w0 + wpl);
        int cnt = 0;
        for (int wd = w0; wd < w1; ++wd) cnt += __popc(s.amask[wd]);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - cnt;
        for (int wd = w0; wd < w1; ++wd) { s.wpre[wd] = run; run += __popc(s.amask[wd]); }
        if (lane == 31) s.misc[M_NB] = incl;
        if (!NS) for (int g = lane; g < RL_MAX_GENES; g += 32) s.cnt_g[g] = 0;
        if (NS) {                                          // ranks of the newborns in row-major order + the serials they take
            int cnb = 0;
            for (int wd = w0; wd < w1; ++wd) cnb += __popc(s.nbmask[wd]);
            int inb = cnb;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xffffffffu, inb, o);
                if (lane >= o) inb += v;
            }
            int rnb = inb - cnb;
            for (int wd = w0; wd < w1; ++wd) { s.nbpre[wd] = rnb; rnb += __popc(s.nbmask[wd]); }
            if (lane == 31) s.misc[M_NNEW] = inb;
        }
    }
    __syncthreads();
    const int nB = s.misc[M_NB];
    rl_agent_rec* rg = P.b.rec + (size_t)w * S;
    for (int c = threadIdx.x; c < C; c += WT) {
        if (ft[c] != RL_AGENT) continue;
        const int slot = s.wpre[c >> 5] + __popc(s.amask[c >> 5] & ((1u << (c & 31)) - 1u));
        const int sc = s.src[c];
        const int h = s.health[sc], g = s.gene[sc];
        const unsigned fl = s.flags[sc] & 0x3Fu;
        atomicAdd(&s.cnt_g[NS ? g : (g & (RL_MAX_GENES - 1))], 1);
        if (slot < S) {
            s.cellof[slot] = (uint16_t)c;
            int4 v;
            v.x = (c & 0xFFFF) | ((int)(uint16_t)(int16_t)h << 16);
            v.y = ((int)(uint16_t)s.age[sc]) | ((int)(uint16_t)s.maxage[sc] << 16);
            v.z = NS ? s.id2gene[g] : g;
            if (NS) {
                const unsigned old = s.aslot[sc];
                double f = 0.0; long long id;
                if (old != RL_NONE16) { f = s.fit[old]; id = s.ser[old]; }
                else id = s.st->next_serial + s.nbpre[c >> 5] + __popc(s.nbmask[c >> 5] & ((1u << (c & 31)) - 1u));
                P.ns.fitness[(size_t)w * S + slot] = f;
                P.ns.serial[(size_t)w * S + slot] = id;
            }
            const unsigned prev = STEP ? (unsigned)s.aslot[sc] : (unsigned)slot;
            v.w = (int)(fl | (((unsigned)(uint8_t)s.action[sc]) << 8) | (prev << 16));
            reinterpret_cast<int4*>(rg)[slot] = v;
            if (STEP)                                   // _get_rewards, environment.py:291-311 (table built above)
            {
                const int ri = g * 8 + ((fl & RL_F_DEAD) ? 2 : 0) + ((fl & RL_F_KILLED) ? 1 : 0);
                P.b.reward[(size_t)w * S + slot] = s.rtab[ri];
                if (P.b.reward_div100) P.b.reward_div100[(size_t)w * S + slot] = s.rtab[ri + 4];
            }
        }
    }
    // ---- padded planes (_prepare_observations :377-404, _get_food :432-446, _get_genes :448-456) ----
    {
        const bool float_path = ft[0] == RL_AGENT;       // np.vectorize dtype quirk of the health plane, SURVEY A.8
        uint32_t* pfg = s.planes; float* ph = reinterpret_cast<float*>(s.planes + PADN);
        const int PH = H + 6;
        const uint32_t magicPW = (uint32_t)(0x100000000ull / (uint64_t)PW) + 1u;
        for (int p = threadIdx.x; p < PH * PW; p += WT) {
            const int pi = (int)__umulhi((uint32_t)p, magicPW), pj = p - pi * PW;
            const int c = s.rowmap[pi] + s.colmap[pj];
            const uint8_t t = ft[c];
            float f = 0.f, hv = -1.f, gv = -2.f;
            if (t == RL_AGENT) {
                const int sc = s.src[c];
                const int h = s.health[sc];
                f = h < 0 ? 1.f : 0.f;
                hv = float_path ? hratio(h) : (float)(h / 200);                  // :396-398
                if (s.flags[sc] & RL_F_DEAD) gv = (float)s.gene[sc];
            } else {
                f = t == RL_FOOD ? .5f : t == RL_SUPER_FOOD ? 1.f : t == RL_POISON ? -1.f : 0.f;
            }
            pfg[p] = (uint32_t)__half_as_ushort(__float2half_rn(f)) | ((uint32_t)__half_as_ushort(__float2half_rn(gv)) << 16);
            ph[p] = hv;
        }
    }
    __syncthreads();
    for (int g = threadIdx.x; g < (NS ? s.misc[M_NIDS] : G); g += WT) {
        int cg = s.cnt_g[g];
        // :357 -- float32(cg / nB); for counts <= 4096 the correctly rounded f32 quotient equals the double-rounded one
        reinterpret_cast<float*>(s.pg)[g] = nB <= 4096 ? __fdiv_rn((float)cg, (float)nB) : (float)((double)cg / (double)nB);
        if (!NS && P.b.gene_count) P.b.gene_count[(size_t)w * G + g] = cg;
        if (NS && cg > 0) atomicAdd(&s.misc[M_NLIN], 1);
    }
    if (threadIdx.x == RL_MAX_GENES) {
        reinterpret_cast<float*>(s.misc)[M_PALL] = (nB <= 4096 && P.cfg.max_agents <= 4096)                     // :358
            ? __fdiv_rn((float)nB, (float)P.cfg.max_agents) : (float)((double)nB / (double)P.cfg.max_agents);
        P.b.n_agents[w] = min(nB, S);
        if (nB > S && P.b.status) atomicOr(&P.b.status[w], 1);
    }
    __syncthreads();
    if (NS) {                                              // hand the world's non-static state back
        if (threadIdx.x == 0) {
            s.st->next_serial += s.misc[M_NNEW];
            if (P.ns.n_lineages) P.ns.n_lineages[w] = s.misc[M_NLIN];
        }
        __syncthreads();
        for (int k = threadIdx.x; k < (int)(sizeof(rl_ns_state) / 8); k += WT)
            reinterpret_cast<long long*>(P.ns.state + w)[k] = reinterpret_cast<const long long*>(s.st)[k];
    }
    if (STEP && P.trace && blockIdx.x == 77 && threadIdx.x == 0) P.trace[6] = clock64();

    // ---- observation rows: one warp per agent ----
    // per-lane constants: element e = lane + 32k of the row -> offset into the padded planes (window elements only)
    if (P.dbg & 2) return;
    __half* obs16 = reinterpret_cast<__half*>(obs_out == P.b.obs_prime ? P.b.obs_prime_h : P.b.obs_state_h);
    int off[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) off[k] = s.offtab[lane + 32 * k];
    const float pall = reinterpret_cast<float*>(s.misc)[M_PALL];
    const int nrow = min(nB, S);
    // Each warp owns a contiguous block of agents and handles it in batches of 16: first lanes 0-15 prepare ONE
    // agent of the batch each (window base, own gene, the six scalars -> per-warp scratch), then the warp emits the
    // rows, broadcasting the per-agent values by shuffle.  Per row: 5 loads, 2 gene decodes, 5 coalesced stores.
    const int per_warp = (nrow + WNW - 1) / WNW;
    const int a0 = warp * per_warp, a1 = min(nrow, a0 + per_warp);
    float* scr = s.ascal + warp * 128;
    const uint32_t* planes = s.planes;
    auto lo_f = [](uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v & 0xFFFFu))); };
    auto hi_f = [](uint32_t v) { return __half2float(__ushort_as_half((unsigned short)(v >> 16))); };
    for (int b0 = a0; b0 < a1; b0 += 16) {
        const int mine = b0 + lane;
        int base_l = 0; float gene_l = 0.f;
        if (lane < 16 && mine < a1) {
            const int d = s.cellof[mine];
            const int i = (int)__umulhi((uint32_t)d, P.magicW), j = d - i * W;
            base_l = i * PW + j;                           // top-left of the 7x7 window in padded coordinates
            const int sc = s.src[d];
            const unsigned fl = s.flags[sc];
            const int g = s.gene[sc];
            gene_l = (float)g;
            float4 lo, hi;
            lo.x = hratio(s.health[sc]);                                        // :365
            lo.y = (fl & RL_F_REPRODUCED) ? 1.f : 0.f;                           // :359
            lo.z = reinterpret_cast<const float*>(s.pg)[g];                     // :357
            lo.w = pall;                                                         // :358
            hi.x = (fl & RL_F_KILLED) ? 1.f : 0.f;                               // :369
            hi.y = (fl & RL_F_ATE_SUPER) ? 1.f : -1.f;                           // :370
            hi.z = 0.f; hi.w = 0.f;
            reinterpret_cast<float4*>(scr + lane * 8)[0] = lo;
            reinterpret_cast<float4*>(scr + lane * 8)[1] = hi;
        }
        __syncwarp();
        const int cnt = min(16, a1 - b0);
        float* orow = obs_out + ((size_t)w * S + b0) * ld;
        uint32_t* orow16 = obs16 ? reinterpret_cast<uint32_t*>(obs16 + ((size_t)w * S + b0) * 160) : nullptr;
        const int sidx = lane >= 19 ? min(lane - 19, 7) : 0;   // lanes 19-24: scalars, 25-31: zero pad (slots 6,7 are zero)
        // float16 copy of a row (320 bytes; element 159 := 1.0): lane l holds elements 32 k + l.  Each lane packs its elements of two
        // blocks into one word, swaps words with its neighbour and picks {mine.lo, other.lo} (even lanes: a pair of block k) or
        // {other.hi, mine.hi} (odd lanes: a pair of block k + 1) with one PRMT -- three packs, three shuffles, three 4-byte stores
        const uint32_t hsel = (lane & 1) ? 0x3276u : 0x5410u;
        const int hidx = ((lane & 1) ? 16 : 0) + (lane >> 1);
        // the row loop is instantiated per (padded row, float16 copy) so that nothing uniform is tested per row
        auto emit = [&](auto wide_c, auto h16_c) {
            constexpr bool WIDE = decltype(wide_c)::value, H16 = decltype(h16_c)::value;
#pragma unroll 2
            for (int a = 0; a < cnt; ++a, orow += ld) {
                const int base = __shfl_sync(0xffffffffu, base_l, a);
                const float mygene = __shfl_sync(0xffffffffu, gene_l, a);
                const uint32_t w0 = planes[base + off[0]];      // e  0..31 : food
                const uint32_t w1 = planes[base + off[1]];      // e 32..48 : food, 49..63: health
                const uint32_t w2 = planes[base + off[2]];      // e 64..95 : health
                const uint32_t w3 = planes[base + off[3]];      // e 96,97  : health, 98..127: dead-agent gene
                const uint32_t w4 = planes[base + off[4]];      // e 128..146: dead-agent gene, 147..: scalars / pad
                const float sv = scr[a * 8 + sidx];
                const float v0 = lo_f(w0);
                const float v1 = lane < 17 ? lo_f(w1) : __uint_as_float(w1);
                const float v2 = __uint_as_float(w2);
                const float q3 = hi_f(w3), q4 = hi_f(w4);
                const float g3 = q3 == -2.f ? 0.f : (q3 == mygene ? 1.f : -1.f);             // :424-428
                const float g4 = q4 == -2.f ? 0.f : (q4 == mygene ? 1.f : -1.f);
                const float v3 = lane >= 2 ? g3 : __uint_as_float(w3);
                const float v4 = lane < 19 ? g4 : sv;
                __stcs(orow + lane, v0); __stcs(orow + 32 + lane, v1); __stcs(orow + 64 + lane, v2);
                __stcs(orow + 96 + lane, v3); __stcs(orow + 128 + lane, v4);
                if (WIDE) for (int e = 160 + lane; e < ld; e += 32) __stcs(orow + e, 0.f);
                if (H16) {
                    const uint32_t p01 = pack_h2(v0, v1), p23 = pack_h2(v2, v3), p4 = pack_h2(lane == 31 ? 1.0f : v4, 0.f);
                    const uint32_t n01 = __shfl_xor_sync(0xffffffffu, p01, 1), n23 = __shfl_xor_sync(0xffffffffu, p23, 1);
                    const uint32_t n4 = __shfl_xor_sync(0xffffffffu, p4, 1);
                    __stcs(orow16 + hidx, __byte_perm(p01, n01, hsel));
                    __stcs(orow16 + 32 + hidx, __byte_perm(p23, n23, hsel));
                    if (!(lane & 1)) __stcs(orow16 + 64 + hidx, __byte_perm(p4, n4, 0x5410u));
                    orow16 += 80;
                }
            }
        };
        using T = std::true_type; using F = std::false_type;
        if (ld > 160) { if (orow16) emit(T{}, T{}); else emit(T{}, F{}); }
        else { if (orow16) emit(F{}, T{}); else emit(F{}, F{}); }
        __syncwarp();
    }
}

// =====================================================================================================
// Environment.step -- environment.py:160-186
// =====================================================================================================
template <bool NS>
__global__ void __launch_bounds__(WT, NS ? 1 : RL_WORLD_MINB) k_world_step(const WParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int w = blockIdx.x;
    const int H = P.cfg.height, W = P.cfg.width, C = H * W, Cw = (C + 31) / 32;
    WS s; ws_carve(s, smem, P.cfg.height, P.cfg.width);
    if (NS) ws_carve_ns(s, smem, P.cfg.height, P.cfg.width, P.cfg.slot_cap);
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));

#define WSTAMP(k) do { if (P.trace && blockIdx.x == 77 && threadIdx.x == 0) P.trace[k] = clock64(); } while (0)
    WSTAMP(0);
    const int n = load_world<true, NS>(P, s, w);
    WSTAMP(1);

    // ---- _attack (environment.py:652-699) in closed form + _prepare_movement (:591-625) ----
    for (int sl = threadIdx.x; sl < n; sl += WT) {
        const int c = s.cellof[sl];
        const int a = s.action[c];
        int h = s.health[c];
        unsigned fl = s.flags[c];
        bool hits = false;
        if (a >= 4 && a <= 7) {
            const int tc = nbr(c, a - 4, W, C, P.magicW);
            if (s.type[tc] == RL_AGENT) {
                hits = true;
                fl |= RL_F_KILLED | (s.gene[tc] == s.gene[c] ? RL_F_INTER_KILLED : RL_F_INTRA_KILLED);   // :696-699
            }
        }
        int L = -1;   // largest order index (= cell index) among agents that hit me
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int nb = nbr(c, (d + 2) & 3, W, C, P.magicW);      // the cell whose direction-d neighbour is c
            if (s.type[nb] == RL_AGENT && s.action[nb] == 4 + d) L = max(L, nb);
        }
        if (hits) h = (L > c) ? 0 : min(200, (L >= 0 ? 0 : h) + 100);   // entities.py:178-185
        else if (L >= 0) h = 0;
        s.health[c] = (int16_t)h;
        s.flags[c] = (uint8_t)fl;
        s.tgt[c] = (uint16_t)((a >= 0 && a <= 3) ? nbr(c, a, W, C, P.magicW) : c);
    }
    __syncthreads();

    WSTAMP(2);
    // ---- _execute_movement conflict fixed point (environment.py:637-644, 717-726) ----
    for (;;) {
        int any = 0;
        for (int sl = threadIdx.x; sl < n; sl += WT) {
            const int c = s.cellof[sl];
            const int t = s.tgt[c];
            if (t != c) {
                int cnt = (s.type[t] == RL_AGENT && s.tgt[t] == t) ? 1 : 0;
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    const int nb = nbr(t, d, W, C, P.magicW);
                    cnt += (s.type[nb] == RL_AGENT && s.tgt[nb] == t) ? 1 : 0;
                }
                if (cnt > 1) { s.flags[c] |= 0x40u; any = 1; }
            }
        }
        any = __syncthreads_or(any);
        if (!any) break;
        for (int sl = threadIdx.x; sl < n; sl += WT) {
            const int c = s.cellof[sl];
            if (s.flags[c] & 0x40u) { s.flags[c] &= ~0x40u; s.tgt[c] = (uint16_t)c; }
        }
        __syncthreads();
    }

    WSTAMP(3);
    // ---- execute: _eat (:701-715) on the pre-move content, vanish rule, new grid ----
    for (int c = threadIdx.x; c < C; c += WT) {
        const uint8_t t = s.type[c];
        if (t == RL_AGENT) {
            const bool stay = s.tgt[c] == c;
            s.ntype[c] = stay ? RL_AGENT : RL_EMPTY;
            s.src[c] = stay ? (uint16_t)c : (uint16_t)RL_NONE16;
        } else {
            s.ntype[c] = t;
            s.src[c] = RL_NONE16;
        }
    }
    for (int sl = threadIdx.x; sl < n; sl += WT) {
        const int c = s.cellof[sl];
        const int t = s.tgt[c];
        if (t != c) {
            const uint8_t tt = s.type[t];
            int h = s.health[c];
            if (tt == RL_FOOD) h = min(200, h + 40);
            else if (tt == RL_POISON) h = min(200, h - 40);
            else if (tt == RL_SUPER_FOOD) {
                h = min(200, h + 40);
                s.maxage[c] = (int16_t)(int)((double)s.maxage[c] * 1.2);   // :714
                s.flags[c] |= RL_F_ATE_SUPER;
            } else if (tt == RL_AGENT && t > c) {
                s.flags[c] |= 0x80u;   // erased by the later leaver's grid[old] = Empty (:780)
            }
            s.health[c] = (int16_t)h;
        }
    }
    __syncthreads();
    for (int sl = threadIdx.x; sl < n; sl += WT) {
        const int c = s.cellof[sl];
        const int t = s.tgt[c];
        if (t != c && !(s.flags[c] & 0x80u)) { s.ntype[t] = RL_AGENT; s.src[t] = (uint16_t)c; }
    }
    // ---- _update_death_status (:789-793) + alive counts for _get_rewards (:295-297), vanished included ----
    for (int sl = threadIdx.x; sl < n; sl += WT) {
        const int c = s.cellof[sl];
        if (s.health[c] <= 0 || s.age[c] == s.maxage[c]) s.flags[c] |= RL_F_DEAD;
        else { atomicAdd(&s.misc[M_ALIVE], 1); atomicAdd(&s.alive_g[s.gene[c]], 1); }
    }
    __syncthreads();

    WSTAMP(4);
    // ---- _add_food (:763-776): counts + empty mask by ballot, placements by warp 0 ----
    {
        int nf = 0, np = 0, ns = 0;
        for (int ch = warp; ch < Cw; ch += WNW) {
            const int c = ch * 32 + lane;
            const uint8_t t = c < C ? s.ntype[c] : (uint8_t)255;
            const unsigned m = __ballot_sync(0xffffffffu, t == RL_EMPTY);
            nf += __popc(__ballot_sync(0xffffffffu, t == RL_FOOD));
            np += __popc(__ballot_sync(0xffffffffu, t == RL_POISON));
            ns += __popc(__ballot_sync(0xffffffffu, t == RL_SUPER_FOOD));
            if (lane == 0) s.mask[ch] = m;
        }
        if (lane == 0) {
            if (nf) atomicAdd(&s.misc[M_NFOOD], nf);
            if (np) atomicAdd(&s.misc[M_NPOISON], np);
            if (ns) atomicAdd(&s.misc[M_NSUPER], ns);
        }
    }
    __syncthreads();
    if (warp == 0) {
        const bool want_food = (double)s.misc[M_NFOOD] <= (double)C / 10.0;
        const bool want_poison = (double)s.misc[M_NPOISON] <= (double)C / 20.0;
        const bool want_super = s.misc[M_NSUPER] == 0;
        // set_random draws the index first and the acceptance second (grid.py:75-77): a rejected placement changes
        // nothing, so lanes 0-6 evaluate the seven acceptance draws in parallel and only accepted slots are scanned.
        bool acc = false;
        if (lane < 7) {
            const bool want = lane < 3 ? want_food : lane < 6 ? want_poison : want_super;
            acc = want && rl_uniform(rl_draw(key, P.t, RL_SITE_FOOD_ACCEPT, lane)) < (lane < 6 ? 0.2 : 1.0);
        }
        unsigned todo = __ballot_sync(0xffffffffu, acc);
        while (todo) {
            const int slot = __ffs(todo) - 1;
            todo &= todo - 1;
            const int cell = warp_place(s.mask, Cw, rl_draw(key, P.t, RL_SITE_FOOD_PLACE, slot));
            if (cell < 0) continue;
            if (lane == 0) s.ntype[cell] = slot < 3 ? RL_FOOD : slot < 6 ? RL_POISON : RL_SUPER_FOOD;
            warp_mask_clear(s.mask, cell);
        }
    }
    __syncthreads();
    WSTAMP(5);
    if (P.dbg & 4) { if (threadIdx.x < C) P.b.type[(size_t)w * C + threadIdx.x] = s.ntype[threadIdx.x]; return; }
    finish_and_observe<true, NS>(P, s, w, s.ntype, P.b.obs_prime, n);
    WSTAMP(9);
}

// =====================================================================================================
// Environment.update_env -- environment.py:188-215 (static families)
// =====================================================================================================
template <bool NS>
__global__ void __launch_bounds__(WT, NS ? 1 : RL_WORLD_MINB) k_world_update(const WParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int w = blockIdx.x;
    const int C = P.cfg.height * P.cfg.width, Cw = (C + 31) / 32, G = P.cfg.n_genes;
    WS s; ws_carve(s, smem, P.cfg.height, P.cfg.width);
    if (NS) ws_carve_ns(s, smem, P.cfg.height, P.cfg.width, P.cfg.slot_cap);
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));

    const int n_list = load_world<false, NS>(P, s, w);   // :210 (frozen for the whole phase)

    if (NS && warp == 0) {
        // _update_best_agents (environment.py:728-739): the listed agent with the highest fitness (np.argmax: first maximum
        // in list order) replaces the best-table entry with the lowest fitness (np.argmin: first minimum) if it is fitter
        // and not already in the table (object identity = serial)
        double bv = -1.0e300; int bi = 0x7fffffff;
        for (int sl = lane; sl < n_list; sl += 32) { const double v = s.fit[sl]; if (v > bv) { bv = v; bi = sl; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            rl_ns_state* st = s.st;
            if (n_list > 0) {
                int mi = 0;
                for (int k = 1; k < 10; ++k) if (st->best[k].fitness < st->best[mi].fitness) mi = k;
                const long long id = s.ser[bi];
                bool present = false;
                for (int k = 0; k < 10; ++k) present |= st->best[k].serial == id;
                if (!present && s.fit[bi] > st->best[mi].fitness) {
                    st->best[mi].serial = id; st->best[mi].fitness = s.fit[bi];
                    st->best[mi].brain = s.id2gene[s.gene[s.cellof[bi]]];
                }
            }
            st->produced_gene = -1; st->produced_src_best = -1; st->produced_src_brain = 0;
        }
        __syncwarp();
    }

    for (int ch = warp; ch < Cw; ch += WNW) {
        const int c = ch * 32 + lane;
        const bool in = c < C;
        const uint8_t t = in ? s.type[c] : (uint8_t)255;
        const bool ag = t == RL_AGENT;
        const unsigned fl = ag ? s.flags[c] : 0u;
        const bool elig = ag && !(fl & (RL_F_DEAD | RL_F_REPRODUCED)) && s.age[c] > 5;   // entities.py:244-248
        const unsigned me = __ballot_sync(0xffffffffu, t == RL_EMPTY);
        const unsigned ma = __ballot_sync(0xffffffffu, elig);
        if (lane == 0) { s.mask[ch] = me; s.amask[ch] = ma; }
        if (ag) { if (!NS) atomicOr(&s.misc[M_PRESENT], 1 << s.gene[c]); s.src[c] = (uint16_t)c; }
    }
    __syncthreads();
    // _reproduce (:488-519), the 0.95 trials: trial index = rank among the eligible parents in row-major order, exactly the order
    // of the reference's short-circuited random.random() calls -- a pure function of the eligible mask, so all warps draw them
    // (one 32-cell word each) and leave the successful parents of word ch in s.wpre[ch] (free until finish_and_observe)
    for (int ch = warp; ch < Cw; ch += WNW) {
        const uint32_t m = s.amask[ch];
        int before = 0;
        for (int i = lane; i < ch; i += 32) before += __popc(s.amask[i]);
        before = __reduce_add_sync(0xffffffffu, before);
        const bool el = (m >> lane) & 1u;
        const uint32_t rank = (uint32_t)before + __popc(m & lanemask_lt());
        const bool succ = el && rl_uniform(rl_draw(key, P.t, RL_SITE_REPRO_TRIAL, rank)) > 0.95;
        const uint32_t sm = __ballot_sync(0xffffffffu, succ);
        if (lane == 0) s.wpre[ch] = sm;
    }
    __syncthreads();
    Placer pl = {0, 0};
    if (warp == 0) pl = placer_init(s.mask, Cw);
    if (warp == 0 && n_list <= P.cfg.max_agents) {
        uint32_t birth = 0;
        // births stay sequential: offspring on a uniformly random empty cell (A.9)
        for (int wb = 0; wb < Cw; wb += 32) {
            const uint32_t mine = wb + lane < Cw ? s.wpre[wb + lane] : 0u;
            unsigned has = __ballot_sync(0xffffffffu, mine != 0u);
            while (has) {
                const int wl = __ffs(has) - 1;
                has &= has - 1;
                uint32_t sm = __shfl_sync(0xffffffffu, mine, wl);
                while (sm) {
                    const int bpos = __ffs(sm) - 1;
                    sm &= sm - 1;
                    const int pc = (wb + wl) * 32 + bpos;
                    const int cell = placer_take(pl, s.mask, Cw, rl_draw(key, P.t, RL_SITE_BIRTH_PLACE, birth));
                    if (cell >= 0) {
                        ++birth;
                        spawn_agent(s, cell, s.gene[pc], 200, 0);
                        if (lane == 0) s.src[cell] = (uint16_t)cell;
                    }
                    if (P.cfg.limit_reproduction && lane == 0) s.flags[pc] |= RL_F_REPRODUCED;   // :518-519
                    __syncwarp();
                }
            }
        }
        // _produce (:521-547)
        if (NS) {
            // _produce, non-static (:541-547): gene = ++max_gene (before the placement can fail), brain = deep copy of a random
            // best agent's brain -- the (new gene, source brain id) event is left in the state for the brain pool
            if (rl_uniform(rl_draw(key, P.t, RL_SITE_PRODUCE_TRIAL, 0)) > 0.95) {
                int id = 0;
                if (lane == 0) {
                    rl_ns_state* st = s.st;
                    st->max_gene += 1;
                    const int k = (int)rl_below(rl_draw(key, P.t, RL_SITE_PRODUCE_GENE, 0), 10u);
                    st->produced_gene = st->max_gene; st->produced_src_best = k; st->produced_src_brain = st->best[k].brain;
                }
                const int cell = placer_take(pl, s.mask, Cw, rl_draw(key, P.t, RL_SITE_BIRTH_PLACE, birth));
                if (cell >= 0) {
                    if (lane == 0) {
                        id = s.misc[M_NIDS];
                        if (id < NS_IDS) { s.id2gene[id] = s.st->max_gene; s.misc[M_NIDS] = id + 1; }
                        else { id = NS_IDS - 1; if (P.b.status) atomicOr(&P.b.status[w], 4); }
                    }
                    id = __shfl_sync(0xffffffffu, id, 0);
                    spawn_agent(s, cell, id, 200, 0);
                    if (lane == 0) s.src[cell] = (uint16_t)cell;
                }
            }
        } else if (rl_uniform(rl_draw(key, P.t, RL_SITE_PRODUCE_TRIAL, 0)) > 0.95) {
            const uint32_t all = G >= 32 ? 0xffffffffu : ((1u << G) - 1u);
            uint32_t cand = all & ~(uint32_t)s.misc[M_PRESENT];
            if (!cand) cand = all;
            const int pick = (int)rl_below(rl_draw(key, P.t, RL_SITE_PRODUCE_GENE, 0), (uint32_t)__popc(cand));
            const int gene = (int)__fns(cand, 0, pick + 1);
            const int cell = placer_take(pl, s.mask, Cw, rl_draw(key, P.t, RL_SITE_BIRTH_PLACE, birth));
            if (cell >= 0) {
                spawn_agent(s, cell, gene, 200, 0);
                if (lane == 0) s.src[cell] = (uint16_t)cell;
            }
        }
    }
    __syncthreads();
    // _remove_dead_agents (:795-799); a dead agent's cell becomes food, so the empty-cell mask (and warp 0's Placer) stay valid.
    // The surviving agents of word ch are left in s.amask[ch] for the top-up's head count.
    for (int ch = warp; ch < Cw; ch += WNW) {
        const int c = ch * 32 + lane;
        const bool ag = c < C && s.type[c] == RL_AGENT;
        const bool dead = ag && (s.flags[c] & RL_F_DEAD);
        if (dead) s.type[c] = RL_FOOD;
        const unsigned m = __ballot_sync(0xffffffffu, ag && !dead);
        if (lane == 0) s.amask[ch] = m;
    }
    __syncthreads();
    if (!NS && P.target > 0) {
        // fused saturated-world generator (rl_world_update_top_up): exactly what k_world_topup does on the state this kernel would
        // have written -- the same draws (keyed by t and the placement counter), one list rebuild and one observation pass instead of two
        if (warp == 0) {
            int cur = 0;
            for (int i = lane; i < Cw; i += 32) cur += __popc(s.amask[i]);
            cur = min(__reduce_add_sync(0xffffffffu, cur), P.cfg.slot_cap);
            for (uint32_t k = 0; cur < P.target; ++k, ++cur) {
                const int cell = placer_take(pl, s.mask, Cw, rl_draw(key, P.t, RL_SITE_TOPUP_PLACE, k));
                if (cell < 0) break;
                const int gene = (int)rl_below(rl_draw(key, P.t, RL_SITE_TOPUP_GENE, k), (uint32_t)P.cfg.n_genes);
                const int health = 10 * (1 + (int)rl_below(rl_draw(key, P.t, RL_SITE_TOPUP_HEALTH, k), 20));
                const int age = (int)rl_below(rl_draw(key, P.t, RL_SITE_TOPUP_AGE, k), (uint32_t)P.max_age);
                spawn_agent(s, cell, gene, health, age);
                if (lane == 0) s.src[cell] = (uint16_t)cell;
            }
        }
        __syncthreads();
    }
    finish_and_observe<false, NS>(P, s, w, s.type, P.b.obs_state, 0);
}

// =====================================================================================================
// Environment.reset -- environment.py:133-158
// =====================================================================================================
template <bool NS>
__global__ void __launch_bounds__(WT) k_world_reset(const WParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int w = blockIdx.x;
    const int C = P.cfg.height * P.cfg.width, Cw = (C + 31) / 32, G = P.cfg.n_genes;
    WS s; ws_carve(s, smem, P.cfg.height, P.cfg.width);
    if (NS) ws_carve_ns(s, smem, P.cfg.height, P.cfg.width, P.cfg.slot_cap);
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));
    for (int c = threadIdx.x; c < C; c += WT) { s.type[c] = RL_EMPTY; s.src[c] = (uint16_t)c; s.aslot[c] = RL_NONE16; }
    for (int k = threadIdx.x; k < M_WORDS; k += WT) s.misc[k] = 0;
    init_tables(P, s);
    if (NS) {
        // the first agents carry gene i = lineage id i; the ten initial best agents are deep copies of agent 0 (:149): serials
        // -1..-10, fitness 0, private brain copies -1..-10; max_gene = len(brains) (:108)
        for (int k = threadIdx.x; k < NS_IDS; k += WT) { s.cnt_g[k] = 0; s.alive_g[k] = 0; s.id2gene[k] = k < G ? k : -1; }
        if (threadIdx.x == 0) {
            rl_ns_state* st = s.st;
            st->max_gene = G; st->produced_gene = -1; st->produced_src_best = -1; st->produced_src_brain = 0; st->next_serial = 0;
            for (int k = 0; k < 10; ++k) { st->best[k].serial = -1 - k; st->best[k].fitness = 0.0; st->best[k].brain = -1 - k; st->best[k]._pad = 0; }
        }
    }
    __syncthreads();
    if (NS && threadIdx.x == 0) s.misc[M_NIDS] = G;
    build_empty_mask(s.type, s.mask, C);
    __syncthreads();
    if (warp == 0) {
        for (int g = 0; g < G; ++g) {                                           // :148
            const int cell = warp_place(s.mask, Cw, rl_draw(key, 0, RL_SITE_RESET_AGENT_PLACE, g));
            if (cell >= 0) { spawn_agent(s, cell, g, 200, 0); warp_mask_clear(s.mask, cell); }
        }
        for (int pass = 0; pass < 2; ++pass) {                                  // _init_food :759-761
            const double p = pass == 0 ? 0.1 : 0.05;
            const uint32_t trial = pass == 0 ? RL_SITE_RESET_FOOD_TRIAL : RL_SITE_RESET_POISON_TRIAL;
            const uint32_t place = pass == 0 ? RL_SITE_RESET_FOOD_PLACE : RL_SITE_RESET_POISON_PLACE;
            uint32_t k = 0;
            for (int base = 0; base < C; base += 32) {
                const int i = base + lane;
                const bool ok = i < C && rl_uniform(rl_draw(key, 0, trial, (uint32_t)i)) < p;
                int succ = __popc(__ballot_sync(0xffffffffu, ok));
                for (; succ > 0; --succ) {
                    const int cell = warp_place(s.mask, Cw, rl_draw(key, 0, place, k++));
                    if (cell >= 0) {
                        if (lane == 0) s.type[cell] = pass == 0 ? RL_FOOD : RL_POISON;
                        warp_mask_clear(s.mask, cell);
                    }
                }
            }
        }
        const int cell = warp_place(s.mask, Cw, rl_draw(key, 0, RL_SITE_RESET_SUPER_PLACE, 0));   // :757
        if (cell >= 0) { if (lane == 0) s.type[cell] = RL_SUPER_FOOD; warp_mask_clear(s.mask, cell); }
    }
    __syncthreads();
    finish_and_observe<false, NS>(P, s, w, s.type, P.b.obs_state, 0);
}

// =====================================================================================================
// saturated-world generator -- harness, mirrored by oracle rlo_topup / RefWorld.top_up
// =====================================================================================================
__global__ void __launch_bounds__(WT, RL_WORLD_MINB) k_world_topup(const WParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int w = blockIdx.x;
    const int C = P.cfg.height * P.cfg.width, Cw = (C + 31) / 32;
    WS s; ws_carve(s, smem, P.cfg.height, P.cfg.width);
    const int warp = threadIdx.x >> 5;
    const uint64_t key = rl_world_key(P.cfg.seed, (uint64_t)(P.cfg.world_id0 + w));
    const int n = load_world<false, false>(P, s, w);
    for (int c = threadIdx.x; c < C; c += WT) s.src[c] = (uint16_t)c;
    build_empty_mask(s.type, s.mask, C);
    __syncthreads();
    if (warp == 0) {
        int cur = n;
        for (uint32_t k = 0; cur < P.target; ++k, ++cur) {
            const int cell = warp_place(s.mask, Cw, rl_draw(key, P.t, RL_SITE_TOPUP_PLACE, k));
            if (cell < 0) break;
            const int gene = (int)rl_below(rl_draw(key, P.t, RL_SITE_TOPUP_GENE, k), (uint32_t)P.cfg.n_genes);
            const int health = 10 * (1 + (int)rl_below(rl_draw(key, P.t, RL_SITE_TOPUP_HEALTH, k), 20));
            const int age = (int)rl_below(rl_draw(key, P.t, RL_SITE_TOPUP_AGE, k), (uint32_t)P.max_age);
            spawn_agent(s, cell, gene, health, age);
            warp_mask_clear(s.mask, cell);
        }
    }
    __syncthreads();
    finish_and_observe<false, false>(P, s, w, s.type, P.b.obs_state, 0);
}

__global__ void __launch_bounds__(WT) k_world_observe(const WParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int w = blockIdx.x;
    const int C = P.cfg.height * P.cfg.width;
    WS s; ws_carve(s, smem, P.cfg.height, P.cfg.width);
    load_world<false, false>(P, s, w);
    for (int c = threadIdx.x; c < C; c += WT) s.src[c] = (uint16_t)c;
    __syncthreads();
    finish_and_observe<false, false>(P, s, w, s.type, P.which ? P.b.obs_prime : P.b.obs_state, 0);
}

size_t g_world_smem_max = 0, g_world_smem_max_ns = 0;

int world_prepare(const rl_world_cfg* cfg, const rl_world_bufs* b, WParams& P, size_t& smem, const rl_world_ns_bufs* ns = nullptr) {
    RL_ARG_CHECK(cfg && b);
    RL_ARG_CHECK(cfg->height >= 3 && cfg->width >= 3);          // World/grid.py:23-24
    RL_ARG_CHECK((int64_t)cfg->height * cfg->width <= 65535);
    RL_ARG_CHECK(cfg->n_worlds > 0 && cfg->n_genes > 0 && cfg->n_genes <= RL_MAX_GENES);
    RL_ARG_CHECK(cfg->slot_cap > 0 && cfg->slot_cap <= cfg->height * cfg->width);
    RL_ARG_CHECK(cfg->obs_ld >= 160 && cfg->obs_ld % 32 == 0 && cfg->obs_ld <= 192);
    RL_ARG_CHECK((!b->obs_state_h && !b->obs_prime_h) || (cfg->obs_ld == 160 && b->obs_state_h && b->obs_prime_h));
    RL_ARG_CHECK(cfg->max_agents > 0);
    if (!cfg->static_families && !ns)
        return rl_set_err(RL_ERR_UNSUPPORTED, "static_families=False goes through rl_world_reset_ns / rl_world_step_ns / rl_world_update_ns");
    if (ns) {
        RL_ARG_CHECK(!cfg->static_families && ns->fitness && ns->serial && ns->state);
        P.ns = *ns;
    } else {
        P.ns.fitness = nullptr; P.ns.serial = nullptr; P.ns.state = nullptr; P.ns.n_lineages = nullptr;
    }
    RL_ARG_CHECK(b->type && b->rec && b->n_agents && b->reward && b->obs_state && b->obs_prime);
    P.trace = nullptr; P.cfg = *cfg; P.b = *b; P.t = 0; P.target = 0; P.max_age = 50; P.which = 0;
    { const char* d = getenv("RL_WORLD_DEBUG"); P.dbg = d ? atoi(d) : 0; }
    P.magicW = (uint32_t)(0x100000000ull / (uint64_t)cfg->width) + 1u;
    smem = ws_bytes(cfg->height, cfg->width);
    if (ns) smem = ((smem + 15) & ~(size_t)15) + ws_bytes_ns(cfg->height, cfg->width, cfg->slot_cap);
    if (smem > 227 * 1024) return rl_set_err(RL_ERR_UNSUPPORTED, "world of %d cells (slot_cap %d) needs %zu B shared memory", cfg->height * cfg->width, cfg->slot_cap, smem);
    if (ns) {
        if (smem > g_world_smem_max_ns) {
            const int v = (int)smem;
            RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_step<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
            RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_update<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
            RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_reset<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
            g_world_smem_max_ns = smem;
        }
        return RL_OK;
    }
    if (smem > g_world_smem_max) {
        const int v = (int)smem;
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_step<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_step<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_update<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_topup, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_update<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_reset<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_topup, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
        RL_CUDA_CHECK(cudaFuncSetAttribute(k_world_observe, cudaFuncAttributeMaxDynamicSharedMemorySize, v));
        g_world_smem_max = smem;
    }
    return RL_OK;
}

}  // namespace

extern "C" {

int rl_world_reset(const rl_world_cfg* cfg, const rl_world_bufs* bufs, void* stream) {
    WParams P; size_t smem;
    int rc = world_prepare(cfg, bufs, P, smem);
    if (rc) return rc;
    k_world_reset<false><<<cfg->n_worlds, WT, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_world_step(const rl_world_cfg* cfg, const rl_world_bufs* bufs, uint64_t t, void* stream) {
    WParams P; size_t smem;
    int rc = world_prepare(cfg, bufs, P, smem);
    if (rc) return rc;
    P.t = t;
    static long long* trace_dev = nullptr;
    const bool tracing = getenv("RL_WORLD_TRACE") != nullptr && cfg->n_worlds > 77;
    if (tracing) {
        if (!trace_dev) RL_CUDA_CHECK(cudaMalloc(&trace_dev, 16 * sizeof(long long)));
        P.trace = trace_dev;
    }
    k_world_step<false><<<cfg->n_worlds, WT, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    if (tracing) {
        long long h[16];
        RL_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
        RL_CUDA_CHECK(cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[world trace] load %lld attack %lld conflict %lld execute+death %lld add_food %lld list+planes %lld observe %lld total %lld\n",
                h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5], h[9] - h[6], h[9] - h[0]);
    }
    return RL_OK;
}

int rl_world_update(const rl_world_cfg* cfg, const rl_world_bufs* bufs, uint64_t t, void* stream) {
    WParams P; size_t smem;
    int rc = world_prepare(cfg, bufs, P, smem);
    if (rc) return rc;
    P.t = t;
    k_world_update<false><<<cfg->n_worlds, WT, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_world_update_top_up(const rl_world_cfg* cfg, const rl_world_bufs* bufs, uint64_t t, int32_t target, int32_t max_age,
                           void* stream) {
    WParams P; size_t smem;
    int rc = world_prepare(cfg, bufs, P, smem);
    if (rc) return rc;
    RL_ARG_CHECK(max_age > 0 && target >= 0);
    P.t = t; P.target = target; P.max_age = max_age;
    k_world_update<false><<<cfg->n_worlds, WT, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_world_top_up(const rl_world_cfg* cfg, const rl_world_bufs* bufs, uint64_t t, int32_t target, int32_t max_age,
                    void* stream) {
    WParams P; size_t smem;
    int rc = world_prepare(cfg, bufs, P, smem);
    if (rc) return rc;
    RL_ARG_CHECK(max_age > 0 && target >= 0);
    P.t = t; P.target = target; P.max_age = max_age;
    k_world_topup<<<cfg->n_worlds, WT, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_world_observe(const rl_world_cfg* cfg, const rl_world_bufs* bufs, int32_t which, void* stream) {
    WParams P; size_t smem;
    int rc = world_prepare(cfg, bufs, P, smem);
    if (rc) return rc;
    P.which = which;
    k_world_observe<<<cfg->n_worlds, WT, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_world_reset_ns(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_world_ns_bufs* ns, void* stream) {
    WParams P; size_t smem;
    RL_ARG_CHECK(ns);
    int rc = world_prepare(cfg, bufs, P, smem, ns);
    if (rc) return rc;
    k_world_reset<true><<<cfg->n_worlds, WT, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_world_step_ns(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_world_ns_bufs* ns, uint64_t t, void* stream) {
    WParams P; size_t smem;
    RL_ARG_CHECK(ns);
    int rc = world_prepare(cfg, bufs, P, smem, ns);
    if (rc) return rc;
    P.t = t;
    k_world_step<true><<<cfg->n_worlds, WT, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

int rl_world_update_ns(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_world_ns_bufs* ns, uint64_t t, void* stream) {
    WParams P; size_t smem;
    RL_ARG_CHECK(ns);
    int rc = world_prepare(cfg, bufs, P, smem, ns);
    if (rc) return rc;
    P.t = t;
    k_world_update<true><<<cfg->n_worlds, WT, smem, (cudaStream_t)stream>>>(P);
    RL_CUDA_CHECK(cudaGetLastError());
    return RL_OK;
}

}  // extern "C"

/*
 * rl_rng.h -- counter-based random draws for the ReinLife hot path (a SPEC, shared by
 * the CUDA kernels, the C oracle and -- restated in numpy/python ints -- the reference shim).
 *
 * The reference draws from three process-global MT19937 streams (python `random`,
 * legacy `np.random`, torch) that the env and the brains interleave
 * (call sites World/grid.py:75-77, World/environment.py:501,528,
 * 536-538,760, Models/PERD3QN.py:165,205,209, Models/DQN.py:135-137).  A single interleaved
 * stream cannot be consumed by thousands of worlds stepping in parallel, so "same seeds" is
 * DEFINED as: every draw is a pure function of (seed, global world id, env step counter t,
 * call site, index-within-site).  The reference is run against the same function by
 * rebinding its module-level `random` / `np` names (oracle/ref_harness.py), so both sides
 * consume identical numbers.
 */
#ifndef RL_RNG_H
#define RL_RNG_H
#include <stdint.h>

#if defined(__CUDACC__)
#define RL_HD __host__ __device__ __forceinline__
#else
#define RL_HD static inline
#endif

/* call-site ids (index meaning in brackets) */
enum rl_rng_site {
    RL_SITE_RESET_AGENT_PLACE  = 1,  /* [gene i]            grid.py:75 via environment.py:148 */
    RL_SITE_RESET_FOOD_TRIAL   = 2,  /* [trial i < H*W]     environment.py:760 (Food, p=.1)   */
    RL_SITE_RESET_FOOD_PLACE   = 3,  /* [k-th success]      grid.py:75 via environment.py:761 */
    RL_SITE_RESET_POISON_TRIAL = 4,  /* [trial i < H*W]     environment.py:760 (Poison, .05)  */
    RL_SITE_RESET_POISON_PLACE = 5,  /* [k-th success]                                        */
    RL_SITE_RESET_SUPER_PLACE  = 6,  /* [0]                 environment.py:757                */
    RL_SITE_FOOD_PLACE         = 7,  /* [slot 0-2 food, 3-5 poison, 6 super] grid.py:75 via environment.py:767-776 */
    RL_SITE_FOOD_ACCEPT        = 8,  /* [same slot]         grid.py:77 (p=.2 / 1)             */
    RL_SITE_REPRO_TRIAL        = 9,  /* [rank among eligible parents] environment.py:501       */
    RL_SITE_BIRTH_PLACE        = 10, /* [k-th placement of this update_env] grid.py:75 via environment.py:515,539 */
    RL_SITE_PRODUCE_TRIAL      = 11, /* [0]                 environment.py:528                */
    RL_SITE_PRODUCE_GENE       = 12, /* [0]                 environment.py:536/538            */
    RL_SITE_TOPUP_PLACE        = 13, /* [k-th top-up agent] harness-only (SURVEY 8d saturated generator) */
    RL_SITE_TOPUP_GENE         = 14,
    RL_SITE_TOPUP_HEALTH       = 15,
    RL_SITE_TOPUP_AGE          = 16,
    RL_SITE_ACT_EXPLORE        = 20, /* [slot]  PERD3QN.py:205 / D3QN.py:168 / DQN.py:135     */
    RL_SITE_ACT_RANDOM         = 21, /* [slot]  PERD3QN.py:209 / D3QN.py:172 / DQN.py:137     */
    RL_SITE_ACT_SAMPLE         = 22, /* [slot]  PPO.py:166-167 (inverse CDF on one uniform)   */
    RL_SITE_REPLAY_SAMPLE      = 30, /* [event_rank*batch + i]  PERD3QN.py:165 (np.random.choice, with replacement) */
    RL_SITE_REPLAY_SAMPLE_UNIFORM = 31,/* [((event_rank*n_iter + iter) << 9) + c]  c-th _randbelow call of random.sample:
                                          D3QN.py:140 (n_iter 1), DQN.py:100 inside the 5-iteration loop of DQN.py:143 */
    RL_SITE_SUMTREE_SAMPLE     = 32  /* [((event_rank*64 + stratum) << 6) + redraw]  random.uniform of PERDQN.py:292 */
};

RL_HD uint64_t rl_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

/* per-world key: computed once per kernel */
RL_HD uint64_t rl_world_key(uint64_t seed, uint64_t world_id) {
    return rl_mix64(seed ^ rl_mix64(world_id + 0x9E3779B97F4A7C15ull));
}

/* the draw: 64 random bits */
RL_HD uint64_t rl_draw(uint64_t world_key, uint64_t step, uint32_t site, uint32_t idx) {
    uint64_t x = rl_mix64(world_key + step * 0xD1342543DE82EF95ull + 0x9E3779B97F4A7C15ull);
    return rl_mix64(x ^ (((uint64_t)site << 32) | (uint64_t)idx));
}

/* uniform double in [0,1): 53 bits, same construction as random.random() */
RL_HD double rl_uniform(uint64_t bits) { return (double)(bits >> 11) * (1.0 / 9007199254740992.0); }

/* integer in [0,n): multiply-high on the top 32 bits (n < 2^32) */
RL_HD uint32_t rl_below(uint64_t bits, uint32_t n) { return (uint32_t)(((bits >> 32) * (uint64_t)n) >> 32); }

#endif /* RL_RNG_H */

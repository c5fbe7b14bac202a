/*
 * reinlife_b200.h -- C ABI of libreinlife_b200.so, the B200 (sm_100a) implementation of
 * ReinLife's data-parallel hot path.  Plain pointers and sizes only; every pointer in a
 * *_bufs struct is a DEVICE pointer unless the function name ends in _host.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream).  Every entry point returns 0
 * on success, a negative rl_status otherwise, and never synchronises the stream unless
 * documented.
 *
 * Each entry point names the reference interface (file:line under the reference checkout)
 * it replaces; INTEGRATION.md shows the ctypes binding a ReinLife maintainer would add.
 */
#ifndef REINLIFE_B200_H
#define REINLIFE_B200_H
#include <stdint.h>
#include "rl_rng.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RL_OBS_DIM 153          /* World/environment.py:119,362-370 */
#define RL_N_ACTIONS 8          /* World/utils.py:4-17 */
#define RL_FOV 3                /* World/environment.py:353 */
#define RL_MAX_GENES 32         /* static families: one gene per brain */

/* World/utils.py:20-33 */
enum rl_entity { RL_EMPTY = 0, RL_FOOD = 1, RL_POISON = 2, RL_AGENT = 3, RL_KIN = 4, RL_SUPER_FOOD = 5 };

/* rl_agent_rec.flags bits (World/entities.py:145-160) */
#define RL_F_KILLED 1u          /* agent.killed */
#define RL_F_INTER_KILLED 2u    /* agent.inter_killed (set when victim has the SAME gene, environment.py:696-697) */
#define RL_F_INTRA_KILLED 4u    /* agent.intra_killed */
#define RL_F_ATE_SUPER 8u       /* agent.ate_super_food == 1.0 (else -1) */
#define RL_F_REPRODUCED 16u     /* agent.reproduced */
#define RL_F_DEAD 32u           /* agent.dead */

enum rl_status {
    RL_OK = 0,
    RL_ERR_ARG = -1,            /* bad size / null pointer (reference: AssertionError, World/grid.py:23-24) */
    RL_ERR_CUDA = -2,           /* a CUDA call failed; rl_last_error() has the text */
    RL_ERR_UNSUPPORTED = -3
};

/* One agent, 16 bytes, lists are kept in ROW-MAJOR cell order = the order of
 * Grid.get_entities (World/grid.py:60-67), which is the reference's agent order everywhere. */
typedef struct rl_agent_rec {
    uint16_t cell;              /* i*width + j */
    int16_t  health;            /* multiple of 10, may be negative (World/environment.py:269,711) */
    int16_t  age;
    int16_t  max_age;
    int32_t  gene;
    uint8_t  flags;
    int8_t   action;            /* 0-3 move U/R/D/L, 4-7 attack U/R/D/L, -1 = none yet */
    uint16_t prev_slot;         /* slot this agent had in the previous `state` list (obs_state row) */
} rl_agent_rec;

typedef struct rl_world_cfg {
    int32_t  n_worlds;          /* worlds in this shard */
    int32_t  height, width;     /* >= 3 (World/grid.py:23-24); height*width <= 65535 */
    int32_t  n_genes;           /* = len(brains) for static families */
    int32_t  max_agents;        /* Environment(max_agents) -- a soft cap */
    int32_t  slot_cap;          /* rows per world in rec/reward/obs buffers (<= height*width) */
    int32_t  obs_ld;            /* floats per observation row, >= 153, multiple of 4; pad is zero-filled */
    int32_t  static_families;   /* 0: the *_ns entry points (rl_world_ns_bufs) */
    int32_t  limit_reproduction;
    int32_t  incentivize_killing;
    uint64_t seed;
    int64_t  world_id0;         /* global id of local world 0 (sharding: rank*n_worlds) */
} rl_world_cfg;

typedef struct rl_world_bufs {
    uint8_t*      type;         /* [n_worlds, H*W] rl_entity per cell */
    rl_agent_rec* rec;          /* [n_worlds, slot_cap] current agent list */
    int32_t*      n_agents;     /* [n_worlds] */
    float*        reward;       /* [n_worlds, slot_cap] float32(reference reward), written by step */
    float*        obs_state;    /* [n_worlds, slot_cap, obs_ld] agent.state  (written by reset/update/top_up) */
    float*        obs_prime;    /* [n_worlds, slot_cap, obs_ld] agent.state_prime (written by step) */
    int32_t*      gene_count;   /* [n_worlds, n_genes] agents per gene in the current list */
    int32_t*      status;       /* [n_worlds] sticky bit 0: slot_cap overflow */
    float*        stats;        /* [n_worlds, n_genes, RL_N_STATS] tracker partials of the last step, may be NULL */
    float*        reward_div100;/* [n_worlds, slot_cap] float32(reference reward / 100.0) -- what PPOAgent.learn stores
                                   (Models/PPO.py:73); written by step, may be NULL (needed only with PPO brains) */
    void*         obs_state_h;  /* optional float16 copies of obs_state / obs_prime ([n_worlds, slot_cap, 160] halves, column 159 = 1.0),
                                   written by the same kernels next to the float32 rows (obs_ld must be 160): what the tensor-core   */
    void*         obs_prime_h;  /* consumers read (TMA row gathers of rl_brain_act_p, float16 replay rows of rl_replay_store); NULL = off */
} rl_world_bufs;

/* per world x gene partial sums over the post-step agent list (Helpers/tracker.py:178-266) */
enum rl_stat { RL_STAT_COUNT = 0, RL_STAT_AGE_SUM, RL_STAT_REWARD_SUM, RL_STAT_AGE_MAX,
               RL_STAT_ATTACKS, RL_STAT_KILLS, RL_N_STATS = 8 };

const char* rl_last_error(void);
int rl_version(void);

/* Environment.reset()  -- World/environment.py:133-158.  Writes type/rec/n_agents/obs_state. */
int rl_world_reset(const rl_world_cfg* cfg, const rl_world_bufs* bufs, void* stream);

/* Environment.step()   -- World/environment.py:160-186 (_act :258, _attack :652, _prepare_movement :591,
 * _execute_movement :627, _eat :701, _update_death_status :789, _get_rewards :277, _add_food :763,
 * _get_observations :313).  Reads rec[].action (set by the act call or by the caller), `t` is the env
 * step counter (1 for the first step after reset).  Writes type/rec/n_agents/reward/obs_prime. */
int rl_world_step(const rl_world_cfg* cfg, const rl_world_bufs* bufs, uint64_t t, void* stream);

/* Environment.update_env() -- World/environment.py:188-215 (_reproduce :488, _produce :521,
 * _remove_dead_agents :795, _get_observations :313, _update_agents_state :784).  Same `t` as the step
 * it follows.  Writes type/rec/n_agents/obs_state. */
int rl_world_update(const rl_world_cfg* cfg, const rl_world_bufs* bufs, uint64_t t, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Non-static families (static_families = False): World/environment.py:149 (ten deep-copied best agents), :506-507
 * (offspring share the parent's brain), :541-547 (_produce: gene = ++max_gene with a deep copy of a random best agent's
 * brain), :728-739 (_update_best_agents on Agent.fitness with object identity).  The World side: per-agent float64
 * fitness and a serial number (object identity, handed out in row-major order to agents that are new at the end of
 * reset / update_env), and per world max_gene, the ten best agents and the _produce event of the last update_env --
 * the event (new gene, source brain id) is what a brain pool executes.  Layouts equal oracle/rl_oracle.c (rlo_ns).
 * A brain id is the gene of a lineage, or -1-k for the private copy of initial best agent k.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rl_ns_best { int64_t serial; double fitness; int32_t brain; int32_t _pad; } rl_ns_best;
typedef struct rl_ns_state {
    int32_t max_gene;            /* Environment.max_gene */
    int32_t produced_gene;       /* gene created by the last update_env, -1 if none (set even when the grid was full) */
    int32_t produced_src_best;   /* index into best[] that random.choice picked (:543) */
    int32_t produced_src_brain;  /* brain id that is deep-copied */
    int64_t next_serial;
    rl_ns_best best[10];
} rl_ns_state;
typedef struct rl_world_ns_bufs {
    double*      fitness;        /* [n_worlds, slot_cap] Agent.fitness of the listed agents */
    int64_t*     serial;         /* [n_worlds, slot_cap] object identity of the listed agents */
    rl_ns_state* state;          /* [n_worlds] */
    int32_t*     n_lineages;     /* [n_worlds] distinct genes in the current list (tracker: Helpers/tracker.py:178-186), may be NULL */
} rl_world_ns_bufs;
/* reset / step / update_env with cfg->static_families = 0 (same contracts as the static entry points; status bit 2 = more
 * than 256 distinct lineages alive in one world). */
int rl_world_reset_ns(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_world_ns_bufs* ns, void* stream);
int rl_world_step_ns(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_world_ns_bufs* ns, uint64_t t, void* stream);
int rl_world_update_ns(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_world_ns_bufs* ns, uint64_t t, void* stream);

/* Saturated-world generator of the benchmark: add agents on random empty cells until
 * `target` are on the grid (gene U{0..G-1}, health 10*U{1..20}, age U{0..max_age-1}), then observe. */
int rl_world_top_up(const rl_world_cfg* cfg, const rl_world_bufs* bufs, uint64_t t, int32_t target,
                    int32_t max_age, void* stream);
/* rl_world_update followed by rl_world_top_up in ONE launch (the benchmark loop's update_env + saturate): the same draws and the same
 * final state, bit for bit (tests/test_world_gpu.py), with one agent-list rebuild and one observation pass instead of two. */
int rl_world_update_top_up(const rl_world_cfg* cfg, const rl_world_bufs* bufs, uint64_t t, int32_t target, int32_t max_age,
                           void* stream);

/* Environment._get_observations() alone -- World/environment.py:313-375, Grid.fov World/grid.py:90-117.
 * Writes obs_state (which=0) or obs_prime (which=1) for the current list; does not change the world. */
int rl_world_observe(const rl_world_cfg* cfg, const rl_world_bufs* bufs, int32_t which, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Row lists: per-brain compaction of the agent lists of all worlds (deterministic: world-major,
 * then slot order).  A row id is  world*slot_cap + slot  and addresses rec / reward / obs_* rows.
 * ---------------------------------------------------------------------------------------------- */
enum rl_row_kind {
    RL_ROWS_ALL = 0,    /* every listed agent of the gene: the get_action loop, Helpers/trainer.py:88-89        */
    RL_ROWS_STORE = 1,  /* age > 1: transitions that reach brain.learn, World/entities.py:194-208             */
    RL_ROWS_EVENT = 2,  /* age > 1 and (age % train_freq == 0 or dead): train() triggers, Models/PERD3QN.py:120-122 */
    RL_N_ROW_KINDS = 3
};

typedef struct rl_rows_bufs {
    int32_t* count;     /* [n_genes*3, n_worlds]  per-world counts                   */
    int32_t* offset;    /* [n_genes*3, n_worlds]  exclusive prefix over worlds       */
    int32_t* total;     /* [n_genes*3]                                               */
    int32_t* rows;      /* [n_genes*3, row_cap]   row ids                            */
    int32_t  row_cap;
    int32_t  _pad;
} rl_rows_bufs;

/* train_freq[g] > 0; event_on[g] = 1 when brain g trains this step (n_epi > exploration), else no EVENT rows.
 * kinds_mask: bit k set = build kind k.  Three tiny kernels (count, scan, scatter); no host sync. */
int rl_rows_build(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows,
                  const int32_t* train_freq_host, const int32_t* event_on_host, int32_t kinds_mask, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Brains.  Parameters live in one flat float32 buffer per network in KERNEL LAYOUT
 * (input-major, i.e. transposed nn.Linear weights; first layer zero-padded from 153 to 160 inputs):
 *
 *   [ W1t 160 x N1 | b1 N1 | W2t N1 x N2 | b2 N2 | Wh N2 x NH | bh NH | pad to 4 | W2 N2 x N1 (output-major copy) ]
 *
 *   RL_MODEL_DUELING (D3QN, PERD3QN; Models/PERD3QN.py:185-202): N1=128, N2=256 = [adv_fc1 | value_fc1], NH=9 =
 *       [adv_fc2 (rows 0-127 only) | value_fc2 (rows 128-255 only)]; the other NH entries are structural zeros.
 *   RL_MODEL_DQN (Models/DQN.py:119-130): N1=128, N2=64, NH=8.
 *   RL_MODEL_PPO (Models/PPO.py:96-112): N1=256, N2=256, NH=9 = [fc_pi | fc_v].
 * The output-major copy of W2 is maintained by rl_brain_adam / rl_brain_sync_w2 for the backward pass.
 * ---------------------------------------------------------------------------------------------- */
enum rl_model_kind { RL_MODEL_DUELING = 0, RL_MODEL_DQN = 1, RL_MODEL_PPO = 2 };

typedef struct rl_model_dims { int32_t n1, n2, nh, off_b1, off_w2t, off_b2, off_wh, off_bh, off_w2, n_train, n_total; } rl_model_dims;
int rl_model_get_dims(int32_t kind, rl_model_dims* out);

/* act-rule per brain (who explores how) */
enum rl_act_rule {
    RL_ACT_DUELING = 0, /* u > eps ? argmax : choice(range(8))      Models/PERD3QN.py:204-210, D3QN.py:167-173 */
    RL_ACT_DQN = 1,     /* coin < eps ? randint(0,7) : argmax        Models/DQN.py:132-139                      */
    RL_ACT_PPO = 2,     /* inverse-CDF sample of softmax(pi)         Models/PPO.py:164-169                      */
    RL_ACT_PERDQN = 3   /* u <= eps ? randrange(8) : argmax          Models/PERDQN.py:101-111 (epsilon moves in train_model) */
};

typedef struct rl_brain_act {
    int32_t kind;           /* rl_model_kind */
    int32_t rule;           /* rl_act_rule   */
    const float* params;    /* online network, kernel layout */
    const double* epsilon;  /* DEVICE scalar: the brain's current epsilon (see rl_brain_epsilon_update) */
} rl_brain_act;

/* epsilon schedules, evaluated on the device so that -- exactly like the reference, where the schedule lives inside
 * get_action -- a brain with no listed agent this step does not advance (no host sync needed to know that):
 *   RL_ACT_DUELING: if training and n_epi > seen: eps *= decay while eps > eps_min; seen = n_epi   (PERD3QN.py:82-86, D3QN.py:84-89)
 *   RL_ACT_DQN:     if training and n_epi % 30 == 0: eps = max(0.01, 0.20 - 0.20*(n_epi/max_epi))  (DQN.py:67-69)
 * eps_dev [n_brains] double, seen_dev [n_brains] int64 are device arrays owned by the caller. */
typedef struct rl_brain_sched {
    int32_t rule;           /* rl_act_rule */
    int32_t training;
    double  eps_min, decay;
    int64_t max_epi;
} rl_brain_sched;
int rl_brain_epsilon_update(const rl_rows_bufs* rows, const rl_brain_sched* sched_host, int32_t n_brains, int64_t n_epi,
                            double* eps_dev, int64_t* seen_dev, void* stream);

/* brain.get_action for every listed agent of every world (Helpers/trainer.py:88-89, Helpers/tester.py:58-68):
 * forward on obs_state rows, exploration draws keyed (t_act, slot), result written to rec[].action.
 * q_out (optional) [n_genes, row_cap, 8]: network outputs per list position (Q values, or pi for PPO);
 * prob_out (optional) [n_worlds*slot_cap]: pi(a) of the sampled action for PPO brains (agent.prob). */
int rl_brain_act_all(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows,
                     const rl_brain_act* brains_host, int32_t n_brains, uint64_t t_act,
                     float* q_out, float* prob_out, void* stream);
/* (a brain whose `kind` is negative is skipped by rl_brain_act_all: it acts through rl_brain_act_tc) */


/* ------------------------------------------------------------------------------------------------
 * Replay + learn.  One ring per (world, brain): N worlds = N independent reference runs that share
 * only the brains' weights (so at N=1 the buffer is exactly Models/PERD3QN.py:133-182).
 * ---------------------------------------------------------------------------------------------- */
typedef struct rl_replay_bufs {             /* per brain; ring index = local world */
    float*   obs;        /* [n_worlds, capacity, obs_ld]  state        (float32, or float16 when obs_fp16 = 1) */
    float*   next_obs;   /* [n_worlds, capacity, obs_ld]  state_prime  (same element type as obs)               */
    int8_t*  action;     /* [n_worlds, capacity] */
    float*   reward;     /* [n_worlds, capacity] float32(reward) (PPO: reward/100, PPO.py:73) */
    uint8_t* done;       /* [n_worlds, capacity] */
    float*   prio;       /* [n_worlds, capacity] raw priorities, zero-initialised (PERD3QN.py:141) */
    float*   pw;         /* [n_worlds, capacity] float32(float64(prio)^0.6) */
    int32_t* len;        /* [n_worlds] */
    int32_t* pos;        /* [n_worlds] */
    int32_t* maxst;      /* [n_worlds][2] or NULL: {float bits of max(priorities), number of entries holding it}, maintained exactly by
                            rl_replay_store / rl_replay_update_prio so that the store needs no scan of the priority array
                            (PERD3QN.py:147 recomputes the max at every memorize); count 0 = unknown: the next store scans.  Zero it
                            after writing prio[] by hand. */
    int32_t  capacity;
    int32_t  prioritized;/* 1: proportional PER (PERD3QN); 0: uniform (D3QN, DQN) */
    int32_t  obs_fp16;   /* 1: obs / next_obs rows are stored as float16 (obs_ld halves per row, column obs_ld-1 = 1.0): the ring
                            of the dueling brains under precision="fp16" -- the tensor-core event kernel consumes fp16 operands, so the
                            rows are rounded once at store time instead of at every gather; half the ring memory and traffic.
                            Readers: rl_replay_store (writes), rl_brain_learn_p, rl_brain_learn (fp32 arithmetic on the rounded rows). */
    int32_t  _pad;
} rl_replay_bufs;

/* brain.memorize for every STORE row of `gene` (PERD3QN.py:91-92,143-155): state = obs_state[prev_slot],
 * state_prime = obs_prime[slot]; new items get max(priorities) (1.0 while empty).  Needs rows kinds STORE. */
int rl_replay_store(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                    const rl_replay_bufs* replay, void* stream);

/* buffer.sample(batch) for every EVENT row of `gene` (PERD3QN.py:157-175): proportional sampling with
 * replacement from the world's ring, exact-integer CDF, draws keyed (t, RL_SITE_REPLAY_SAMPLE,
 * event_rank*batch+i).  sample_idx: [row_cap, batch] int32, one row per event in EVENT-list order. */
int rl_replay_sample(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                     int32_t batch, uint64_t t, int32_t* sample_idx, void* stream);

/* random.sample(buffer, batch) for every EVENT row of `gene` (D3QN.py:138-142, DQN.py:99-100): uniform, WITHOUT
 * replacement -- CPython's Random.sample restated on the counter RNG (pool method for len <= 277, set method with
 * redraws above; population index counts from the oldest item of the deque).  `iter`/`n_iter` select the draw stream
 * of the DQN 5-iteration loop (DQN.py:143); events whose ring holds <= min_len items (DQN.py:79: size() > 1000) or
 * fewer than `batch` items (where the reference raises ValueError: bit 0 of *status is set, status may be NULL) get
 * sample_idx[...] = -1 and are skipped by the learn kernels. */
int rl_replay_sample_uniform(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                             int32_t batch, uint64_t t, int32_t iter, int32_t n_iter, int32_t min_len, int32_t* sample_idx,
                             int32_t* status, void* stream);

/* buffer.update_priorities (PERD3QN.py:177-179) for all events of `gene`, in event order, later writes win. */
int rl_replay_update_prio(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                          int32_t batch, const int32_t* sample_idx, const float* new_prio, void* stream);

typedef struct rl_learn_bufs {              /* per brain */
    float*   params;        /* eval / online network, kernel layout [n_total] */
    float*   target;        /* target network [n_total] */
    float*   grad_scratch;  /* [n_cta, n_train] per-CTA partial sums (n_cta = rl_learn_grid()) */
    float*   grad;          /* [n_train + 4]: summed gradient; grad[n_train] = number of events summed */
    float*   adam_m;        /* [n_train] */
    float*   adam_v;        /* [n_train] */
    float*   mask;          /* [n_train] 1 = trainable, 0 = padding / structural zero */
    int32_t* adam_step;     /* [1] device step counter (advances only when events > 0) */
    float*   new_prio;      /* [row_cap, batch] |max_a Q_target(s') - Q(s,a)| per sampled row */
    float*   loss;          /* [row_cap] per-event MSE loss */
    int32_t  kind;          /* rl_model_kind */
    int32_t  batch;         /* 64 */
    float    gamma;
    float    lr;
} rl_learn_bufs;

int rl_learn_grid(void);    /* CTAs used by rl_brain_learn = rows of grad_scratch */

/* One train() event per EVENT row of `gene` (PERD3QN.py:94-115 / D3QN.py:97-116): two forwards, MSE TD loss,
 * explicit backward.  Per-event gradients are SUMMED into `grad` (grad[n_train] = #events); call
 * rl_brain_adam afterwards (after an optional all-reduce of `grad` across ranks). */
int rl_brain_learn(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                   const int32_t* sample_idx, const rl_learn_bufs* learn, void* stream);

/* One of the 5 iterations of train(q, q_target, memory, optimizer) (Models/DQN.py:142-153) for every EVENT row of `gene`:
 * q_a = q(s)[a], y = r + gamma * max_a q_target(s') * done_mask, smooth-L1 (mean over the 32 sampled rows), explicit
 * backward.  learn->kind = RL_MODEL_DQN, learn->batch = 32; sample_idx [row_cap, 32] from rl_replay_sample_uniform
 * (events marked -1 are skipped).  Gradients are summed over events into `grad`, grad[n_train] = number of events that
 * trained; follow with rl_brain_adam -- five (sample, learn, adam) rounds make one train() call. */
int rl_brain_learn_dqn(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                       const int32_t* sample_idx, const rl_learn_bufs* learn, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PERDQN (Models/PERDQN.py).  Network 153-64-64-8 (:311-323) = RL_MODEL_DQN layout with the first hidden layer
 * zero-padded from 64 to 128 units (padded weights have mask 0 and stay exactly 0).  Memory (:262-308) = the brain's
 * rl_replay_bufs ring (prioritized = 0; pos = SumTree.write, len = n_entries) plus one SumTree per world:
 * Sequence per learn step:
 *     rl_sumtree_add -> rl_replay_store -> rl_sumtree_sample -> rl_brain_learn_perdqn -> [all-reduce grad] ->
 *     rl_brain_adam -> rl_sumtree_update -> rl_perdqn_epsilon_step -> rl_brain_sync_target(cond = #EVENT rows)
 * ---------------------------------------------------------------------------------------------- */
typedef struct rl_sumtree_bufs {            /* per brain */
    double*  tree;          /* [n_worlds, 2*capacity-1] float64 nodes, zero-initialised; leaf of slot d = d+capacity-1 (:198-259) */
    double*  beta;          /* [n_worlds] Memory.beta, initialised to 0.4, +0.001 per sample() up to 1 (:266-267,284) */
    int32_t* status;        /* [1] sticky bit 0: a stratum needed more than 64 redraws (:290-295); may be NULL */
    int32_t  capacity;      /* = replay->capacity */
    int32_t  train_start;   /* train_model only once n_entries >= train_start (1000, :69,192) */
    float    p_new;         /* priority of a new item = float32(0.01)**0.6: append_sample's error is always 0 (see
                               csrc/sumtree_kernels.cu) */
    int32_t  _pad;
} rl_sumtree_bufs;

/* Memory.add (:275-277, SumTree.add :229-240) for every STORE row of `gene`, agent order, float32 propagation of the
 * reference's tensor path.  Call BEFORE rl_replay_store of the same step (it reads the ring's write position). */
int rl_sumtree_add(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                   const rl_sumtree_bufs* tree, void* stream);

/* Memory.sample(64) (:278-303) for every EVENT row of `gene`: stratified draws keyed (t, RL_SITE_SUMTREE_SAMPLE, ...),
 * redraw while the leaf holds no data.  sample_idx [row_cap, 64] = data slots (-1 for events of a world whose memory
 * holds < train_start items: no train_model call, :192); ev_weight [row_cap] = mean of float32(is_weight) of the event,
 * the factor of loss = (is_weights * mse_loss(pred, target)).mean() (:182). */
int rl_sumtree_sample(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                      const rl_sumtree_bufs* tree, int32_t batch, uint64_t t, int32_t* sample_idx, float* ev_weight,
                      void* stream);

/* train_model's forward/backward (:144-186) for every EVENT row: pred = model(s)[a], target = r + (1-done) * gamma *
 * max target_model(s'), loss as above; errors |pred - target| -> learn->new_prio [row_cap, 64]; gradients summed over
 * events into learn->grad, grad[n_train] = events that trained.  learn->kind = RL_MODEL_DQN, learn->batch = 64. */
int rl_brain_learn_perdqn(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                          const int32_t* sample_idx, const float* ev_weight, const rl_learn_bufs* learn, void* stream);
/* rl_brain_learn_dqn / rl_brain_learn_perdqn on the tensor cores (csrc/tc_dqn_kernels.cu: 128-row batch-major tiles, fp16 operands /
 * fp32 accumulation, both nets' operand images resident in shared memory, all weight gradients resident in TMEM).  Same contract and
 * outputs at the tolerance of the dueling tensor-core kernels; the Environment's choice for DQN / PERDQN under precision="fp16". */
int rl_brain_learn_dqn_p(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                         const int32_t* sample_idx, const rl_learn_bufs* learn, void* stream);
int rl_brain_learn_perdqn_p(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                            const int32_t* sample_idx, const float* ev_weight, const rl_learn_bufs* learn, void* stream);
/* get_action of the DQN-layout brains (DQN, PERDQN) on the tensor cores (csrc/tc_dqn_kernels.cu::k_act_dqn_p): as rl_brain_act_all for
 * one gene, operand images built in the kernel from brain->params. */
int rl_brain_act_dqn_p(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                       const rl_brain_act* brain, uint64_t t_act, float* q_out, void* stream);

/* Memory.update (:305-308) for the 64 sampled leaves of every trained event: event order, batch order, duplicates
 * included, priority = powf(|error| + 0.01, 0.6), float64 propagation. */
int rl_sumtree_update(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                      const rl_sumtree_bufs* tree, int32_t batch, const int32_t* sample_idx, const float* errors,
                      void* stream);

/* train_model's `if epsilon > epsilon_min: epsilon -= epsilon_decay` (:132-133), applied once per optimizer step that
 * happened (learn->grad[n_train] > 0); eps_dev = the brain's device epsilon (rl_brain_act.epsilon). */
int rl_perdqn_epsilon_step(const rl_learn_bufs* learn, double* eps_dev, double eps_min, double eps_decay, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PPO (Models/PPO.py:62-77,113-162).  One python-list-like data buffer per (world, brain): `traj` (an rl_replay_bufs
 * used append-only: len = items held, pos unused, prio[] = pi_old(a), reward[] = reward/100).  A step's train triggers
 * (age % train_freq == 0 or dead, PPO.py:75) cut the list into segments, one learn() call each; rows after the last
 * trigger stay for the next step.  Sequence per learn step:
 *     rl_ppo_store -> k_epoch x ( rl_ppo_epoch -> [all-reduce grad] -> rl_brain_adam ) -> rl_ppo_compact
 * ---------------------------------------------------------------------------------------------- */
typedef struct rl_ppo_bufs {                /* per brain */
    rl_replay_bufs traj;     /* capacity = rows per world (<= 8192) */
    uint8_t* seg_end;        /* [n_worlds, capacity] 1 = this transition's agent triggered learn() */
    int32_t* n_cons;         /* [n_worlds] rows consumed by this step's learn() calls */
    int32_t* row_off;        /* [n_worlds + 1] exclusive prefix of n_cons; row_off[n_worlds] = rows of this step */
    int32_t* flat_src;       /* [row_cap] world*capacity + j of every consumed row, world-major, list order */
    int32_t* row_T;          /* [row_cap] length of the row's segment (the batch of its learn() call) */
    uint8_t* row_end;        /* [row_cap] 1 = last row of its segment */
    float*   td;             /* [row_cap] td_target  (PPO.py:140) */
    float*   delta;          /* [row_cap] td_target - v(s)  (:141) */
    float*   adv;            /* [row_cap] GAE advantage (:143-150) */
    int32_t* status;         /* [1] sticky: bit 0 = data list overflowed `capacity`, bit 1 = row_cap overflow; may be NULL */
    int32_t  row_cap;
    float    lmbda;          /* 0.95 */
    float    eps_clip;       /* 0.1  */
    int32_t  _pad;
} rl_ppo_bufs;

/* put_data for every STORE row of `gene` in agent order (state = obs_state[prev_slot], prob = prob[w*slot_cap + prev_slot]
 * as written by rl_brain_act_all, reward = bufs->reward_div100), then the segment plan of this step (3 kernels). */
int rl_ppo_store(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                 const float* prob, int32_t train_freq, const rl_ppo_bufs* ppo, void* stream);

/* One of the k_epoch optimizer steps of PPO.learn() for every segment: v(s), v(s') -> td_target, delta; GAE per segment
 * (float32, list order); clipped surrogate + smooth-L1 value loss, explicit backward.  learn->kind = RL_MODEL_PPO,
 * learn->gamma = 0.98.  Per-segment gradients are summed into `grad`, grad[n_train] = number of segments. */
int rl_ppo_epoch(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_ppo_bufs* ppo,
                 const rl_learn_bufs* learn, void* stream);

/* self.data = [] for the consumed rows (PPO.py:133): the rows after the last trigger move to the front. */
int rl_ppo_compact(const rl_world_cfg* cfg, int32_t gene, const rl_ppo_bufs* ppo, void* stream);

/* torch.optim.Adam defaults on grad/grad[n_train] (mean over events); no-op when no event happened.
 * Also refreshes the output-major copy of W2. */
int rl_brain_adam(const rl_learn_bufs* learn, void* stream);

/* target_net.load_state_dict(eval_net.state_dict())  (PERD3QN.py:124-125).  cond (device int32, may be NULL): copy only
 * when *cond > 0 -- the reference syncs inside learn(), i.e. only if some agent of the brain called learn this step. */
int rl_brain_sync_target(const rl_learn_bufs* learn, const int32_t* cond, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Tracker partials (Helpers/tracker.py:178-266): per gene, summed over all worlds of the shard, for the
 * current agent list (call after rl_world_step: that is the list tracker.update_results sees,
 * World/environment.py:206-207).  Deterministic two-level reduction, no host sync.
 * out (device, float64): [n_genes][RL_N_STATS] then 8 trailing doubles:
 *   [0] total agents, [1] worlds with >= 1 agent, [2] sum over worlds of #distinct genes present,
 *   [3] ctrl[0] echoed (step stamp written by the host into `ctrl`), [4..7] reserved.
 * RL_STAT_COUNT = agents, AGE_SUM, REWARD_SUM (float64 sum of float32 rewards), AGE_MAX, ATTACKS (action >= 4),
 * KILLS (sum of agent.killed), [6] = worlds in which the gene is present, [7] reserved.
 * ---------------------------------------------------------------------------------------------- */
int rl_world_stats(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const int64_t* ctrl_dev, double* scratch_dev,
                   int32_t* counter_dev, double* out_dev, void* stream);
int rl_world_stats_scratch_doubles(const rl_world_cfg* cfg);   /* size of scratch_dev in doubles */


/* Test hook: one-tile tcgen05 kind::tf32 GEMM  D[M,N] = A * B^T  on interleaved no-swizzle operand images
 * (csrc/tc_tile.cuh).  a_mn / b_mn = 1: the operand image is stored k-rows x mn-columns (MN-major). */
/* Timing probe (scripts/tc_mma_bench.py): average cycles to issue / to complete `ksteps` back-to-back tcgen05.mma
 * kind::tf32 (K = 8 each) for operand layout `mode` (0 no-swizzle, 1 no-swizzle with padded chunk stride, 2 SWIZZLE_128B). */
int rl_tc_mma_bench(int M, int N, int ksteps, int mode, int iters, long long* out_host);
/* Timing probe (scripts/tc_issue_probe.py): cycles to issue / complete `nmma` back-to-back M = 128 tcgen05.mma kind::f16 of
 * width N from `n_warps` warps; layout 0 K-major no-swizzle, 1 MN-major no-swizzle, 2 K-major SWIZZLE_128B; style 0 = issuing
 * thread picked by `lane == 0` (descriptors in vector registers), style 1 = warp-uniform loop + elect.sync (uniform registers). */
int rl_tc_issue_probe(int N, int nmma, int layout, int style, int n_warps, int iters, long long* out_host);
/* Test hook: rl_tc_gemm_test_h with a swizzle layout type per operand (descriptor bits 61-63: 0 none, 2 SWIZZLE_128B) and, when
 * a_kblk / b_kblk != 0, k-step addresses of the form (ks / 4) * kblk + (ks % 4) * kstep (four K = 16 steps per 128-byte atom). */
int rl_tc_gemm_test_hx(const void* a_img, const void* b_img, float* d, int M, int N, int K, int a_halves, int b_halves,
                       uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                       int a_mn, int b_mn, uint32_t a_kblk, uint32_t b_kblk, int a_layout, int b_layout, void* stream);
/* Test hook: TMA tile::gather4 of rows idx[0..127] of the fp16 tensor [n_rows][160] at `ring` into three [128 rows][128 B]
 * SWIZZLE_128B K blocks; `out` receives the 48 KB shared-memory image (tests/test_tc_gpu.py pins the layout). */
int rl_tma_gather_test(const void* ring, long long n_rows, const int32_t* idx, void* out, int box_rows);

int rl_tc_gemm_test(const float* a_img, const float* b_img, float* d, int M, int N, int K, int a_mn, int b_mn, void* stream);
/* Probe hook behind it (scripts/tc_probe.py): the same tf32 tile GEMM with caller-supplied B descriptor geometry
 * (leading / stride byte offsets, byte advance per k-step; 0 = the defaults of rl_tc_gemm_test). */
int rl_tc_gemm_test_ex(const float* a_img, const float* b_img, float* d, int M, int N, int K, int a_mn, int b_mn,
                       int b_lbo, int b_sbo, int b_kstep, void* stream);
/* Test hook: one-tile tcgen05 kind::f16 GEMM  D[M,N] = A * B^T  on fp16 operand images with caller-supplied descriptor
 * geometry (leading / stride byte offsets, byte advance per K = 16 step) and major-ness (a_mn / b_mn = 1: MN-major). */
int rl_tc_gemm_test_h(const void* a_img, const void* b_img, float* d, int M, int N, int K, int a_halves, int b_halves,
                      uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_kstep,
                      int a_mn, int b_mn, void* stream);


/* ------------------------------------------------------------------------------------------------
 * Tensor-core path (tcgen05.mma kind::tf32, fp32 accumulation in TMEM) for the dueling brains.
 * Same contract and outputs as rl_brain_learn; network products use 10-bit-mantissa operands
 * (stated tolerance: tests/test_tc_gpu.py).  Weight images (`wimg`, rl_tc_wimg_floats() floats per network)
 * are derived from the kernel-layout parameters with rl_brain_build_wimg after every parameter change.
 * ---------------------------------------------------------------------------------------------- */
int rl_tc_wimg_floats(void);
/* brain.get_action of ONE dueling brain on the tensor cores (same rule, draws and outputs as rl_brain_act_all; the
 * network products use tf32 operands, so Q values agree with the fp32 path to the tolerance of tests/test_tc_gpu.py). */
int rl_brain_act_tc(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                    const rl_brain_act* brain, const float* wimg_eval, uint64_t t_act, float* q_out, void* stream);
int rl_brain_build_wimg(int32_t kind, const float* params, float* wimg, void* stream);
int rl_brain_learn_tc(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                      const int32_t* sample_idx, const rl_learn_bufs* learn, const float* wimg_eval, const float* wimg_target,
                      void* stream);

/* fp16-operand variant of the tensor-core event kernel (tcgen05.mma kind::f16, fp32 accumulation; fp16 carries the same 11
 * significant bits as tf32, K = 16 per instruction, half the operand bytes).  `wimg_*_h` are fp16 weight images of
 * rl_tc_wimg_floats() HALVES each, built by rl_brain_build_wimg_h.  Same contract and outputs as rl_brain_learn_tc. */
int rl_brain_build_wimg_h(int32_t kind, const float* params, void* wimg_h, void* stream);
/* get_action of ONE dueling brain with fp16 operands (same contract as rl_brain_act_tc, fp16 weight image) */
int rl_brain_act_h(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                   const rl_brain_act* brain, const void* wimg_eval_h, uint64_t t_act, float* q_out, void* stream);
/* get_action of the dueling brains in the batch-major form of rl_brain_learn_p (csrc/tc_act_kernels.cu: 128-row tiles, one warp-uniform
 * MMA issuer, rows gathered float32 -> fp16 into a SWIZZLE_128B image by dedicated warps).  Same contract as rl_brain_act_h; the
 * Environment's default under precision="fp16". */
int rl_brain_act_p(const rl_world_cfg* cfg, const rl_world_bufs* bufs, const rl_rows_bufs* rows, int32_t gene,
                   const rl_brain_act* brain, const void* wimg_eval_h, uint64_t t_act, float* q_out, void* stream);
int rl_brain_learn_h(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                     const int32_t* sample_idx, const rl_learn_bufs* learn, const void* wimg_eval_h, const void* wimg_target_h,
                     void* stream);

/* The same events, TWO per CTA iteration in batch-major form (csrc/tc_pair_kernels.cu: M = 128 tiles, N = 128 / 256 per MMA,
 * MN-major fp16 operands for the weight-gradient GEMMs, no transposed images).  Same contract, weight images, outputs and
 * tolerance as rl_brain_learn_h; the default event kernel of precision="fp16". */
int rl_brain_learn_p(const rl_world_cfg* cfg, const rl_rows_bufs* rows, int32_t gene, const rl_replay_bufs* replay,
                     const int32_t* sample_idx, const rl_learn_bufs* learn, const void* wimg_eval_h, const void* wimg_target_h,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REINLIFE_B200_H */

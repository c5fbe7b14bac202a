/*
 * rl_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C, sequential, CPU restatement of ReinLife's World hot path
 * (reference: ReinLife/World/environment.py, grid.py, entities.py).  It deliberately keeps the
 * reference's *order-dependent, object-by-object* formulation (sequential attack loop, sequential
 * eat/move with grid overwrites, np.unique-style conflict fixed point, list rebuilds by row-major
 * scans) rather than the closed-form parallel rules the CUDA kernels use, so that the two are
 * independent statements of the same semantics.
 *
 * Pinned: oracle/make_golden.py runs the unmodified reference (RNG rebound to include/rl_rng.h,
 * see oracle/ref_harness.py) and tests/test_oracle_golden.py checks this file bit-for-bit against
 * those vectors (tests/golden/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/reinlife_b200.h"

typedef struct {
    int32_t height, width, n_genes, max_agents;
    int32_t static_families, limit_reproduction, incentivize_killing, _pad;
    uint64_t seed;
} rlo_cfg;

typedef struct {
    int i, j, it, jt;
    int health, age, max_age, gene, action;
    int killed, inter_killed, intra_killed, ate_super, reproduced, dead;
    int prev_slot;
    double reward;
    double fitness;    /* Agent.fitness, entities.py:170,189 (python int 0, then float64 sums) */
    int64_t serial;    /* object identity (`agent not in self.best_agents`, environment.py:738); -1 = not numbered yet */
} Agent;

/* ---- non-static families (static_families=False): environment.py:149,506-507,541-547,728-739 ----
 * What the World needs beyond the static state: Agent.fitness, object identity (a serial number, handed out in
 * row-major order to the agents that are new at the end of reset / update_env / top-up), max_gene, and the ten
 * best_agents entries.  A brain is identified by the gene of its lineage (offspring share the parent's brain object,
 * :506-507; every _produce creates gene max_gene+1 with a deepcopy of a best agent's brain, :542-545); the ten initial
 * best agents are deep copies of agent 0 made at reset (:149) and own private copies of brain 0: brain ids -1..-10. */
typedef struct { int64_t serial; double fitness; int32_t brain; int32_t _pad; } rlo_best;
typedef struct {
    int32_t max_gene;            /* Environment.max_gene */
    int32_t produced_gene;       /* gene created by the last update_env, -1 if none (set even when the grid was full) */
    int32_t produced_src_best;   /* index into best[] that random.choice picked (:543) */
    int32_t produced_src_brain;  /* brain id that was deep-copied */
    int64_t next_serial;
    rlo_best best[10];
} rlo_ns;

typedef struct {
    const rlo_cfg* cfg;
    int H, W, C;
    uint8_t* type;     /* [C] entity type, the reference's object grid */
    int* who;          /* [C] agent id when type == AGENT */
    Agent* pool;       /* agent objects */
    int n_pool;
    int* list;         /* self.agents: ids */
    int n_list;
    uint64_t key;
    uint64_t t;
} World;

static void world_init(World* w, const rlo_cfg* cfg, int64_t world_id, uint64_t t) {
    w->cfg = cfg; w->H = cfg->height; w->W = cfg->width; w->C = w->H * w->W;
    w->type = (uint8_t*)calloc(w->C, 1);
    w->who = (int*)malloc(sizeof(int) * w->C);
    w->pool = (Agent*)calloc(2 * w->C + 8, sizeof(Agent));
    w->list = (int*)malloc(sizeof(int) * (2 * w->C + 8));
    w->n_pool = 0; w->n_list = 0;
    w->key = rl_world_key(cfg->seed, (uint64_t)world_id);
    w->t = t;
    for (int c = 0; c < w->C; ++c) w->who[c] = -1;
}
static void world_free(World* w) { free(w->type); free(w->who); free(w->pool); free(w->list); }

static int new_agent(World* w, int cell, int gene) {   /* entities.py:145-160 */
    Agent* a = &w->pool[w->n_pool];
    memset(a, 0, sizeof(*a));
    a->i = cell / w->W; a->j = cell % w->W; a->it = a->i; a->jt = a->j;
    a->health = 200; a->age = 0; a->max_age = 50; a->gene = gene; a->action = -1;
    a->prev_slot = 0xFFFF;
    a->fitness = 0.0; a->serial = -1;
    w->type[cell] = RL_AGENT; w->who[cell] = w->n_pool;
    return w->n_pool++;
}

static void load_state(World* w, const uint8_t* type, const rl_agent_rec* rec, int n) {
    memcpy(w->type, type, w->C);
    for (int s = 0; s < n; ++s) {
        int id = new_agent(w, rec[s].cell, rec[s].gene);
        Agent* a = &w->pool[id];
        a->health = rec[s].health; a->age = rec[s].age; a->max_age = rec[s].max_age;
        a->action = rec[s].action;
        a->killed = !!(rec[s].flags & RL_F_KILLED); a->inter_killed = !!(rec[s].flags & RL_F_INTER_KILLED);
        a->intra_killed = !!(rec[s].flags & RL_F_INTRA_KILLED); a->ate_super = !!(rec[s].flags & RL_F_ATE_SUPER);
        a->reproduced = !!(rec[s].flags & RL_F_REPRODUCED); a->dead = !!(rec[s].flags & RL_F_DEAD);
        a->prev_slot = s;
    }
}

static void load_extra(World* w, const double* fit, const int64_t* ser, int n) {
    if (!fit) return;
    for (int s = 0; s < n; ++s) { w->pool[s].fitness = fit[s]; w->pool[s].serial = ser[s]; }   /* ids = slots after load_state */
}

/* Grid.get_entities(agent): row-major scan -- grid.py:60-67 */
static void rebuild_list(World* w) {
    w->n_list = 0;
    for (int c = 0; c < w->C; ++c)
        if (w->type[c] == RL_AGENT) w->list[w->n_list++] = w->who[c];
}

/* Grid.set_random -- grid.py:69-83.  Returns the chosen cell or -1. `accept_bits` unused when p >= 1. */
static int set_random_cell(World* w, uint64_t place_bits, int use_accept, uint64_t accept_bits, double p) {
    int n_empty = 0;
    for (int c = 0; c < w->C; ++c) n_empty += (w->type[c] == RL_EMPTY);
    if (n_empty == 0) return -1;                       /* ValueError path: no draw consumed */
    int k = (int)rl_below(place_bits, (uint32_t)n_empty);
    int cell = -1;
    for (int c = 0; c < w->C; ++c)
        if (w->type[c] == RL_EMPTY && k-- == 0) { cell = c; break; }
    if (use_accept && !(rl_uniform(accept_bits) < p)) return -1;
    return cell;
}


static void neighbour(const World* w, int i, int j, int dir, int* oi, int* oj) {
    /* up = i-1, right = j+1, down = i+1, left = j-1, toroidal -- environment.py:601-623,664-689 */
    *oi = i; *oj = j;
    if (dir == 0) *oi = (i == 0) ? w->H - 1 : i - 1;
    else if (dir == 1) *oj = (j == w->W - 1) ? 0 : j + 1;
    else if (dir == 2) *oi = (i == w->H - 1) ? 0 : i + 1;
    else *oj = (j == 0) ? w->W - 1 : j - 1;
}

static int imin(int a, int b) { return a < b ? a : b; }

/* ---- Environment._get_observations -- environment.py:313-456, grid.py:90-117 ---- */
static void observe(World* w, double* obs /* [n_list,153] */) {
    rebuild_list(w);                                                       /* :349 */
    int C = w->C, H = w->H, W = w->W;
    double* food = (double*)malloc(sizeof(double) * C);
    double* healthf = (double*)malloc(sizeof(double) * C);
    long* genes = (long*)malloc(sizeof(long) * C);
    /* np.vectorize takes the output dtype from the first element, cell (0,0) -- SURVEY A.8 */
    int float_path = (w->type[0] == RL_AGENT);
    for (int c = 0; c < C; ++c) {
        int tt = w->type[c];
        const Agent* a = tt == RL_AGENT ? &w->pool[w->who[c]] : NULL;
        food[c] = tt == RL_FOOD ? .5 : tt == RL_SUPER_FOOD ? 1. : tt == RL_POISON ? -1. : (a && a->health < 0) ? 1. : 0.;  /* :432-446 */
        double hv = a ? (double)a->health / 200.0 : -1.0;                 /* :396-398 */
        healthf[c] = float_path ? hv : (double)(long)hv;                  /* int64 cast truncates toward zero */
        genes[c] = (a && a->dead) ? a->gene : -2;                         /* :448-456 */
    }
    for (int s = 0; s < w->n_list; ++s) {
        const Agent* a = &w->pool[w->list[s]];
        double* o = obs + (size_t)s * RL_OBS_DIM;
        int same = 0;
        for (int q = 0; q < w->n_list; ++q) same += (w->pool[w->list[q]].gene == a->gene);
        for (int di = -3; di <= 3; ++di)
            for (int dj = -3; dj <= 3; ++dj) {
                int ci = ((a->i + di) % H + H) % H, cj = ((a->j + dj) % W + W) % W;   /* grid.py:99-115 */
                int c = ci * W + cj, e = (di + 3) * 7 + (dj + 3);
                o[e] = food[c];
                o[49 + e] = healthf[c];
                long g = genes[c];                                        /* :424-428 */
                if (g > -1 && g != a->gene) g = -1;
                if (g == a->gene) g = 1;
                if (g == -2) g = 0;
                o[98 + e] = (double)g;
            }
        o[147] = (double)a->health / 200.0;                              /* :365 */
        o[148] = a->reproduced ? 1.0 : 0.0;
        o[149] = (double)same / (double)w->n_list;                        /* :357 */
        o[150] = (double)w->n_list / (double)w->cfg->max_agents;          /* :358 */
        o[151] = (double)a->killed;
        o[152] = a->ate_super ? 1.0 : -1.0;
    }
    free(food); free(healthf); free(genes);
}

/* new agents get their serial in row-major order of the final list (mirrors the harness, which numbers objects at dump time) */
static void store_extra(World* w, double* fit, int64_t* ser, rlo_ns* ns) {
    if (!fit) return;
    rebuild_list(w);
    for (int s = 0; s < w->n_list; ++s) {
        Agent* a = &w->pool[w->list[s]];
        if (a->serial < 0) a->serial = ns->next_serial++;
        fit[s] = a->fitness; ser[s] = a->serial;
    }
}

static void store_state(World* w, uint8_t* type, rl_agent_rec* rec, int32_t* n, double* reward) {
    rebuild_list(w);
    memcpy(type, w->type, w->C);
    *n = w->n_list;
    for (int s = 0; s < w->n_list; ++s) {
        const Agent* a = &w->pool[w->list[s]];
        rl_agent_rec r;
        r.cell = (uint16_t)(a->i * w->W + a->j); r.health = (int16_t)a->health; r.age = (int16_t)a->age;
        r.max_age = (int16_t)a->max_age; r.gene = a->gene;
        r.flags = (uint8_t)((a->killed ? RL_F_KILLED : 0) | (a->inter_killed ? RL_F_INTER_KILLED : 0) |
                            (a->intra_killed ? RL_F_INTRA_KILLED : 0) | (a->ate_super ? RL_F_ATE_SUPER : 0) |
                            (a->reproduced ? RL_F_REPRODUCED : 0) | (a->dead ? RL_F_DEAD : 0));
        r.action = (int8_t)a->action; r.prev_slot = (uint16_t)a->prev_slot;
        rec[s] = r;
        if (reward) reward[s] = a->reward;
    }
}

/* ---- Environment.reset -- environment.py:133-158 ---- */
int rlo_reset(const rlo_cfg* cfg, int64_t world_id, uint8_t* type, rl_agent_rec* rec, int32_t* n, double* obs) {
    World w; world_init(&w, cfg, world_id, 0);
    for (int g = 0; g < cfg->n_genes; ++g) {                               /* :148 */
        int cell = set_random_cell(&w, rl_draw(w.key, 0, RL_SITE_RESET_AGENT_PLACE, (uint32_t)g), 0, 0, 1.0);
        if (cell >= 0) new_agent(&w, cell, g);
    }
    for (int pass = 0; pass < 2; ++pass) {                                 /* _init_food :759-761 */
        double p = pass == 0 ? 0.1 : 0.05;
        uint32_t trial = pass == 0 ? RL_SITE_RESET_FOOD_TRIAL : RL_SITE_RESET_POISON_TRIAL;
        uint32_t place = pass == 0 ? RL_SITE_RESET_FOOD_PLACE : RL_SITE_RESET_POISON_PLACE;
        uint32_t k = 0;
        for (int i = 0; i < w.C; ++i)
            if (rl_uniform(rl_draw(w.key, 0, trial, (uint32_t)i)) < p) {
                int cell = set_random_cell(&w, rl_draw(w.key, 0, place, k++), 0, 0, 1.0);
                if (cell >= 0) w.type[cell] = pass == 0 ? RL_FOOD : RL_POISON;
            }
    }
    {
        int cell = set_random_cell(&w, rl_draw(w.key, 0, RL_SITE_RESET_SUPER_PLACE, 0), 0, 0, 1.0);   /* :757 */
        if (cell >= 0) w.type[cell] = RL_SUPER_FOOD;
    }
    observe(&w, obs);
    for (int s = 0; s < w.n_list; ++s) w.pool[w.list[s]].prev_slot = s;
    store_state(&w, type, rec, n, NULL);
    world_free(&w);
    return 0;
}

/* ---- Environment.step -- environment.py:160-186 ---- */
static int step_impl(const rlo_cfg* cfg, int64_t world_id, uint64_t t, uint8_t* type, rl_agent_rec* rec, int32_t* n,
                     double* reward, double* obs, double* fit, int64_t* ser, rlo_ns* ns) {
    World w; world_init(&w, cfg, world_id, t);
    load_state(&w, type, rec, *n);
    load_extra(&w, fit, ser, *n);
    const int H = w.H, W = w.W;
    (void)H;

    /* _act :258-275 */
    rebuild_list(&w);                                                      /* :267 */
    for (int s = 0; s < w.n_list; ++s) {
        Agent* a = &w.pool[w.list[s]];
        a->health = imin(200, a->health - 10);                             /* :269 */
        a->age = imin(a->max_age, a->age + 1);                             /* :270 */
        a->killed = a->inter_killed = a->intra_killed = 0;                 /* :271 */
    }
    /* _attack :652-699, sequential, pre-move positions */
    for (int s = 0; s < w.n_list; ++s) {
        Agent* a = &w.pool[w.list[s]];
        if (a->dead || a->action < 4 || a->action > 7) continue;
        int ti, tj; neighbour(&w, a->i, a->j, a->action - 4, &ti, &tj);
        int tc = ti * W + tj;
        if (w.type[tc] == RL_AGENT) {                                      /* :692 */
            Agent* v = &w.pool[w.who[tc]];
            v->health = 0;                                                 /* is_attacked, entities.py:183-185 */
            a->health = imin(200, a->health + 100); a->killed = 1;         /* execute_attack, entities.py:178-181 */
            if (v->gene == a->gene) a->inter_killed = 1; else a->intra_killed = 1;   /* :696-699 */
        }
    }
    /* _prepare_movement :591-625 */
    for (int s = 0; s < w.n_list; ++s) {
        Agent* a = &w.pool[w.list[s]];
        if (a->action <= 3 && a->action >= 0 && !a->dead) neighbour(&w, a->i, a->j, a->action, &a->it, &a->jt);
        else { a->it = a->i; a->jt = a->j; }
    }
    /* _execute_movement :627-650 with _get_impossible_coordinates :717-726 */
    {
        int* count = (int*)malloc(sizeof(int) * w.C);
        for (;;) {
            memset(count, 0, sizeof(int) * w.C);
            for (int s = 0; s < w.n_list; ++s) { const Agent* a = &w.pool[w.list[s]]; count[a->it * W + a->jt]++; }
            int any = 0;
            for (int c = 0; c < w.C; ++c) any |= (count[c] > 1);
            if (!any) break;
            for (int s = 0; s < w.n_list; ++s) {
                Agent* a = &w.pool[w.list[s]];
                if (count[a->it * W + a->jt] > 1) { a->it = a->i; a->jt = a->j; }
            }
        }
        free(count);
    }
    for (int s = 0; s < w.n_list; ++s) {                                   /* :647-650 */
        Agent* a = &w.pool[w.list[s]];
        if (a->action > 3) continue;
        int tc = a->it * W + a->jt;
        int tt = w.type[tc];                                               /* _eat :701-715: CURRENT content */
        if (tt == RL_FOOD) a->health = imin(200, a->health + 40);
        else if (tt == RL_POISON) a->health = imin(200, a->health - 40);
        else if (tt == RL_SUPER_FOOD) {
            a->health = imin(200, a->health + 40);
            a->max_age = (int)((double)a->max_age * 1.2);                  /* :714 */
            a->ate_super = 1;
        }
        int oc = a->i * W + a->j;                                          /* _update_agent_position :778-782 */
        w.type[oc] = RL_EMPTY; w.who[oc] = -1;
        w.type[tc] = RL_AGENT; w.who[tc] = w.list[s];
        a->i = a->it; a->j = a->jt;
    }
    /* _update_death_status :789-793 */
    for (int s = 0; s < w.n_list; ++s) {
        Agent* a = &w.pool[w.list[s]];
        if (a->health <= 0 || a->age == a->max_age) a->dead = 1;
    }
    /* _get_rewards :277-311 over the _act list (vanished agents included) */
    for (int s = 0; s < w.n_list; ++s) {
        Agent* a = &w.pool[w.list[s]];
        int kin = 0, alive = 0;
        for (int q = 0; q < w.n_list; ++q) {
            const Agent* o = &w.pool[w.list[q]];
            if (!o->dead) { alive++; if (o->gene == a->gene) kin++; }
        }
        kin = kin - 1 > 0 ? kin - 1 : 0;
        double r;
        if (a->dead) r = (double)(-alive + kin);
        else if (alive == 1) r = 0.0;
        else r = (double)kin / (double)alive;
        if (a->killed && cfg->incentivize_killing) r += 0.2;
        a->reward = r;
        a->fitness += r;                                                   /* update_rl_stats, entities.py:187-192 */
        if (ns)                                                            /* best_agents holds the live object */
            for (int k = 0; k < 10; ++k) if (ns->best[k].serial == a->serial) ns->best[k].fitness = a->fitness;
    }
    /* _add_food :763-776 */
    {
        int nf = 0;
        for (int c = 0; c < w.C; ++c) nf += (w.type[c] == RL_FOOD);
        if ((double)nf <= (double)w.C / 10.0)
            for (uint32_t i = 0; i < 3; ++i) {
                int cell = set_random_cell(&w, rl_draw(w.key, t, RL_SITE_FOOD_PLACE, i), 1, rl_draw(w.key, t, RL_SITE_FOOD_ACCEPT, i), 0.2);
                if (cell >= 0) w.type[cell] = RL_FOOD;
            }
        int np_ = 0;
        for (int c = 0; c < w.C; ++c) np_ += (w.type[c] == RL_POISON);
        if ((double)np_ <= (double)w.C / 20.0)
            for (uint32_t i = 0; i < 3; ++i) {
                int cell = set_random_cell(&w, rl_draw(w.key, t, RL_SITE_FOOD_PLACE, 3 + i), 1, rl_draw(w.key, t, RL_SITE_FOOD_ACCEPT, 3 + i), 0.2);
                if (cell >= 0) w.type[cell] = RL_POISON;
            }
        int ns = 0;
        for (int c = 0; c < w.C; ++c) ns += (w.type[c] == RL_SUPER_FOOD);
        if (ns == 0) {
            int cell = set_random_cell(&w, rl_draw(w.key, t, RL_SITE_FOOD_PLACE, 6), 1, rl_draw(w.key, t, RL_SITE_FOOD_ACCEPT, 6), 1.0);
            if (cell >= 0) w.type[cell] = RL_SUPER_FOOD;
        }
    }
    observe(&w, obs);                                                      /* :186 */
    store_state(&w, type, rec, n, reward);
    store_extra(&w, fit, ser, ns);
    world_free(&w);
    return 0;
}

int rlo_step(const rlo_cfg* cfg, int64_t world_id, uint64_t t, uint8_t* type, rl_agent_rec* rec, int32_t* n,
             double* reward, double* obs) {
    return step_impl(cfg, world_id, t, type, rec, n, reward, obs, NULL, NULL, NULL);
}
int rlo_step_ns(const rlo_cfg* cfg, int64_t world_id, uint64_t t, uint8_t* type, rl_agent_rec* rec, int32_t* n,
                double* reward, double* obs, double* fitness, int64_t* serial, rlo_ns* ns) {
    return step_impl(cfg, world_id, t, type, rec, n, reward, obs, fitness, serial, ns);
}

/* ---- Environment.update_env -- environment.py:188-215 (tracker / best-agents excluded) ---- */
static int update_impl(const rlo_cfg* cfg, int64_t world_id, uint64_t t, uint8_t* type, rl_agent_rec* rec, int32_t* n, double* obs,
                       double* fit, int64_t* ser, rlo_ns* ns) {
    World w; world_init(&w, cfg, world_id, t);
    load_state(&w, type, rec, *n);
    load_extra(&w, fit, ser, *n);
    rebuild_list(&w);                                                      /* :210 (== the list of :349, self.agents) */
    const int n_list = w.n_list;                                           /* frozen for the loop */
    if (ns) {                                                              /* _update_best_agents :728-739 */
        int mi = 0;
        for (int k = 1; k < 10; ++k) if (ns->best[k].fitness < ns->best[mi].fitness) mi = k;      /* np.argmin: first minimum */
        if (n_list > 0) {
            int xi = 0;
            for (int s = 1; s < n_list; ++s) if (w.pool[w.list[s]].fitness > w.pool[w.list[xi]].fitness) xi = s;   /* np.argmax: first maximum */
            const Agent* a = &w.pool[w.list[xi]];
            int present = 0;
            for (int k = 0; k < 10; ++k) present |= (ns->best[k].serial == a->serial);
            if (!present && a->fitness > ns->best[mi].fitness) {
                ns->best[mi].serial = a->serial; ns->best[mi].fitness = a->fitness; ns->best[mi].brain = a->gene;
            }
        }
        ns->produced_gene = -1; ns->produced_src_best = -1; ns->produced_src_brain = 0;
    }
    uint32_t trial = 0, birth = 0;
    /* _reproduce :488-519 */
    for (int s = 0; s < n_list; ++s) {
        Agent* a = &w.pool[w.list[s]];
        int can = (!a->dead && !a->reproduced && a->age > 5);              /* entities.py:244-248 */
        if (can && n_list <= cfg->max_agents &&
            rl_uniform(rl_draw(w.key, t, RL_SITE_REPRO_TRIAL, trial++)) > 0.95) {
            int cell = set_random_cell(&w, rl_draw(w.key, t, RL_SITE_BIRTH_PLACE, birth), 0, 0, 1.0);   /* :515 (A.9) */
            if (cell >= 0) { birth++; new_agent(&w, cell, a->gene); a = &w.pool[w.list[s]]; }
            if (cfg->limit_reproduction) a->reproduced = 1;                /* :518-519 */
        }
    }
    /* _produce :521-547 */
    if (ns) {                                                              /* non-static: :541-547 */
        if (n_list <= cfg->max_agents && rl_uniform(rl_draw(w.key, t, RL_SITE_PRODUCE_TRIAL, 0)) > 0.95) {
            ns->max_gene += 1;                                             /* :542, before the placement can fail */
            const int k = (int)rl_below(rl_draw(w.key, t, RL_SITE_PRODUCE_GENE, 0), 10);   /* random.choice(self.best_agents) :543 */
            ns->produced_gene = ns->max_gene; ns->produced_src_best = k; ns->produced_src_brain = ns->best[k].brain;
            int cell = set_random_cell(&w, rl_draw(w.key, t, RL_SITE_BIRTH_PLACE, birth), 0, 0, 1.0);
            if (cell >= 0) { birth++; new_agent(&w, cell, ns->max_gene); }
        }
    } else if (n_list <= cfg->max_agents && rl_uniform(rl_draw(w.key, t, RL_SITE_PRODUCE_TRIAL, 0)) > 0.95) {
        int present[RL_MAX_GENES]; memset(present, 0, sizeof(present));
        for (int s = 0; s < n_list; ++s) present[w.pool[w.list[s]].gene] = 1;
        int cand[RL_MAX_GENES], nc = 0;
        for (int g = 0; g < cfg->n_genes; ++g) if (!present[g]) cand[nc++] = g;   /* ascending */
        if (nc == 0) for (int g = 0; g < cfg->n_genes; ++g) cand[nc++] = g;
        int gene = cand[rl_below(rl_draw(w.key, t, RL_SITE_PRODUCE_GENE, 0), (uint32_t)nc)];
        int cell = set_random_cell(&w, rl_draw(w.key, t, RL_SITE_BIRTH_PLACE, birth), 0, 0, 1.0);
        if (cell >= 0) { birth++; new_agent(&w, cell, gene); }
    }
    /* _remove_dead_agents :795-799 */
    for (int s = 0; s < n_list; ++s) {
        Agent* a = &w.pool[w.list[s]];
        if (a->dead) { int c = a->i * w.W + a->j; w.type[c] = RL_FOOD; w.who[c] = -1; }
    }
    observe(&w, obs);                                                      /* :214 */
    for (int s = 0; s < w.n_list; ++s) w.pool[w.list[s]].prev_slot = s;   /* state <- state_prime :215 */
    store_state(&w, type, rec, n, NULL);
    store_extra(&w, fit, ser, ns);
    world_free(&w);
    return 0;
}

int rlo_update(const rlo_cfg* cfg, int64_t world_id, uint64_t t, uint8_t* type, rl_agent_rec* rec, int32_t* n, double* obs) {
    return update_impl(cfg, world_id, t, type, rec, n, obs, NULL, NULL, NULL);
}
int rlo_update_ns(const rlo_cfg* cfg, int64_t world_id, uint64_t t, uint8_t* type, rl_agent_rec* rec, int32_t* n, double* obs,
                  double* fitness, int64_t* serial, rlo_ns* ns) {
    return update_impl(cfg, world_id, t, type, rec, n, obs, fitness, serial, ns);
}

/* reset with static_families=False: same world as the static reset (:147-148 does not branch); plus the ten deep copies
 * of agent 0 (:149) and the numbering of the first agents */
int rlo_reset_ns(const rlo_cfg* cfg, int64_t world_id, uint8_t* type, rl_agent_rec* rec, int32_t* n, double* obs,
                 double* fitness, int64_t* serial, rlo_ns* ns) {
    int rc = rlo_reset(cfg, world_id, type, rec, n, obs);
    if (rc) return rc;
    memset(ns, 0, sizeof(*ns));
    ns->max_gene = cfg->n_genes;                                           /* :108 */
    ns->produced_gene = -1; ns->produced_src_best = -1;
    for (int s = 0; s < *n; ++s) { fitness[s] = 0.0; serial[s] = ns->next_serial++; }
    for (int k = 0; k < 10; ++k) { ns->best[k].serial = -1 - k; ns->best[k].fitness = 0.0; ns->best[k].brain = -1 - k; }
    return 0;
}

/* ---- saturated-world generator (harness of SURVEY 8d; mirrored by RefWorld.top_up) ---- */
int rlo_topup(const rlo_cfg* cfg, int64_t world_id, uint64_t t, int32_t target, int32_t max_age,
              uint8_t* type, rl_agent_rec* rec, int32_t* n, double* obs) {
    World w; world_init(&w, cfg, world_id, t);
    load_state(&w, type, rec, *n);
    rebuild_list(&w);
    int cur = w.n_list;
    for (uint32_t k = 0; cur < target; ++k, ++cur) {
        int cell = set_random_cell(&w, rl_draw(w.key, t, RL_SITE_TOPUP_PLACE, k), 0, 0, 1.0);
        if (cell < 0) break;
        int id = new_agent(&w, cell, (int)rl_below(rl_draw(w.key, t, RL_SITE_TOPUP_GENE, k), (uint32_t)cfg->n_genes));
        w.pool[id].health = 10 * (1 + (int)rl_below(rl_draw(w.key, t, RL_SITE_TOPUP_HEALTH, k), 20));
        w.pool[id].age = (int)rl_below(rl_draw(w.key, t, RL_SITE_TOPUP_AGE, k), (uint32_t)max_age);
    }
    observe(&w, obs);
    for (int s = 0; s < w.n_list; ++s) w.pool[w.list[s]].prev_slot = s;
    store_state(&w, type, rec, n, NULL);
    world_free(&w);
    return 0;
}

int rlo_observe(const rlo_cfg* cfg, const uint8_t* type, const rl_agent_rec* rec, int32_t n, double* obs) {
    World w; world_init(&w, cfg, 0, 0);
    load_state(&w, type, rec, n);
    observe(&w, obs);
    world_free(&w);
    return 0;
}

/* Batched drivers: worlds laid out exactly like rl_world_bufs on the host (slot_cap rows per world). */
int rlo_step_many(const rlo_cfg* cfg, int64_t world_id0, int32_t n_worlds, int32_t slot_cap, uint64_t t,
                  uint8_t* type, rl_agent_rec* rec, int32_t* n, double* reward, double* obs) {
    size_t C = (size_t)cfg->height * cfg->width;
    for (int32_t w = 0; w < n_worlds; ++w)
        rlo_step(cfg, world_id0 + w, t, type + w * C, rec + (size_t)w * slot_cap, n + w,
                 reward + (size_t)w * slot_cap, obs + (size_t)w * slot_cap * RL_OBS_DIM);
    return 0;
}
int rlo_update_many(const rlo_cfg* cfg, int64_t world_id0, int32_t n_worlds, int32_t slot_cap, uint64_t t,
                    uint8_t* type, rl_agent_rec* rec, int32_t* n, double* obs) {
    size_t C = (size_t)cfg->height * cfg->width;
    for (int32_t w = 0; w < n_worlds; ++w)
        rlo_update(cfg, world_id0 + w, t, type + w * C, rec + (size_t)w * slot_cap, n + w,
                   obs + (size_t)w * slot_cap * RL_OBS_DIM);
    return 0;
}
int rlo_reset_many(const rlo_cfg* cfg, int64_t world_id0, int32_t n_worlds, int32_t slot_cap,
                   uint8_t* type, rl_agent_rec* rec, int32_t* n, double* obs) {
    size_t C = (size_t)cfg->height * cfg->width;
    for (int32_t w = 0; w < n_worlds; ++w)
        rlo_reset(cfg, world_id0 + w, type + w * C, rec + (size_t)w * slot_cap, n + w,
                  obs + (size_t)w * slot_cap * RL_OBS_DIM);
    return 0;
}
int rlo_topup_many(const rlo_cfg* cfg, int64_t world_id0, int32_t n_worlds, int32_t slot_cap, uint64_t t,
                   int32_t target, int32_t max_age, uint8_t* type, rl_agent_rec* rec, int32_t* n, double* obs) {
    size_t C = (size_t)cfg->height * cfg->width;
    for (int32_t w = 0; w < n_worlds; ++w)
        rlo_topup(cfg, world_id0 + w, t, target, max_age, type + w * C, rec + (size_t)w * slot_cap, n + w,
                  obs + (size_t)w * slot_cap * RL_OBS_DIM);
    return 0;
}

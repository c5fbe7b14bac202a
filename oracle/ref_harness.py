"""TEST INFRASTRUCTURE -- not product code.  Nothing under reinlife_b200/ may import this.

Harness that runs the UNMODIFIED reference (MaartenGr/ReinLife, mounted read-only, default
/root/reference or $REINLIFE_REF) so that golden vectors can be minted from it:

* stubs `pygame` / `matplotlib` (not installed; only imported by the reference's renderer,
  tracker plots and saver -- Helpers/render.py:3, Helpers/tracker.py:2-3, Helpers/saver.py:7),
* rebinds the module-level names `random` and `np` inside ReinLife.World.environment and
  ReinLife.World.grid (and the brain modules) to proxies that serve every draw from the
  counter-based generator specified in include/rl_rng.h, keyed by
  (seed, world, step, call site, index).  The reference source is not edited; only the
  names it looks up at call time are rebound (SURVEY.md Appendix B recommends exactly this),
* dumps the object grid into the canonical SoA form used by the oracle and the CUDA path.

It cannot travel to the GPU box (the reference is absent there); its outputs do, as
tests/golden/*.npz written by oracle/make_golden.py.
"""
import os
import sys
import types

import numpy as _np

REF_PATH = os.environ.get("REINLIFE_REF", "/root/reference")

# ---------------------------------------------------------------- rl_rng.h restated in python ints
M64 = (1 << 64) - 1
SITE = dict(RESET_AGENT_PLACE=1, RESET_FOOD_TRIAL=2, RESET_FOOD_PLACE=3, RESET_POISON_TRIAL=4,
            RESET_POISON_PLACE=5, RESET_SUPER_PLACE=6, FOOD_PLACE=7, FOOD_ACCEPT=8, REPRO_TRIAL=9,
            BIRTH_PLACE=10, PRODUCE_TRIAL=11, PRODUCE_GENE=12, TOPUP_PLACE=13, TOPUP_GENE=14,
            TOPUP_HEALTH=15, TOPUP_AGE=16, ACT_EXPLORE=20, ACT_RANDOM=21, ACT_SAMPLE=22,
            REPLAY_SAMPLE=30, REPLAY_SAMPLE_UNIFORM=31, SUMTREE_SAMPLE=32)


def mix64(z):
    z &= M64
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z


def world_key(seed, world_id):
    return mix64(seed ^ mix64((world_id + 0x9E3779B97F4A7C15) & M64))


def draw(key, step, site, idx):
    x = mix64((key + step * 0xD1342543DE82EF95 + 0x9E3779B97F4A7C15) & M64)
    return mix64(x ^ ((site << 32) | idx))


def uniform(bits):
    return (bits >> 11) * (1.0 / 9007199254740992.0)


def below(bits, n):
    return ((bits >> 32) * n) >> 32


# ---------------------------------------------------------------- draw context
class Ctx:
    """Which (seed, world, step) the next reference call belongs to + per-site sequence counters."""

    def __init__(self):
        self.seed = 0
        self.world = 0
        self.t = 0
        self.key = world_key(0, 0)
        self.counters = {}
        self.pending_food_slot = None
        self.slot = 0          # agent slot during the act phase
        self.event_rank = 0    # replay-sample event rank within (world, brain)
        self.sample_i = 0
        self.last_choice = -1  # index random.choice picked in _produce

    def set_world(self, seed, world):
        self.seed, self.world = seed, world
        self.key = world_key(seed, world)

    def begin(self, t):
        self.t = t
        self.counters = {}
        self.pending_food_slot = None

    def next(self, name):
        v = self.counters.get(name, 0)
        self.counters[name] = v + 1
        return v

    def bits(self, site, idx):
        return draw(self.key, self.t, SITE[site], idx)


CTX = Ctx()


def _caller(depth):
    return sys._getframe(depth + 1)


class _EnvRandom:
    """Stands in for the `random` module inside ReinLife.World.environment."""

    def random(self):
        name = _caller(1).f_code.co_name
        if name == "_reproduce":                    # environment.py:501
            return uniform(CTX.bits("REPRO_TRIAL", CTX.next("repro")))
        if name == "_produce":                      # environment.py:528
            return uniform(CTX.bits("PRODUCE_TRIAL", 0))
        raise RuntimeError(f"unexpected random.random() caller {name}")

    def choice(self, seq):                          # environment.py:536/538 (static), :543 (non-static: best_agents)
        seq = list(seq)
        CTX.last_choice = below(CTX.bits("PRODUCE_GENE", 0), len(seq))
        return seq[CTX.last_choice]

    def randint(self, a, b):                        # environment.py:512 -- dead code
        raise RuntimeError("random.randint reached: _get_empty_within_fov returned coordinates")


class _NpRandom:
    """Stands in for `np.random` inside ReinLife.World.grid / .environment."""

    def randint(self, lo, hi):                      # grid.py:75
        if hi <= lo:
            raise ValueError("low >= high")         # what numpy raises; set_random catches it
        n = hi - lo
        f_set = _caller(1)
        assert f_set.f_code.co_name == "set_random"
        entity = f_set.f_locals["entity"].__name__
        up = _caller(2).f_code.co_name
        CTX.pending_food_slot = None
        if up == "_add_food":                       # environment.py:767-776
            slot = 6 if entity == "SuperFood" else ({"Food": 0, "Poison": 3}[entity] + _caller(2).f_locals["i"])
            CTX.pending_food_slot = slot
            return lo + below(CTX.bits("FOOD_PLACE", slot), n)
        if up == "_init_food":                      # environment.py:757,761
            if entity == "SuperFood":
                return lo + below(CTX.bits("RESET_SUPER_PLACE", 0), n)
            site = "RESET_FOOD_PLACE" if entity == "Food" else "RESET_POISON_PLACE"
            return lo + below(CTX.bits(site, CTX.next(site)), n)
        if up == "_add_agent":
            up2 = _caller(3).f_code.co_name
            if up2 in ("reset", "<listcomp>"):                      # environment.py:148
                return lo + below(CTX.bits("RESET_AGENT_PLACE", CTX.next("reset_agent")), n)
            if up2 in ("_reproduce", "_produce"):   # environment.py:515,539,545
                return lo + below(CTX.bits("BIRTH_PLACE", CTX.next("birth")), n)
        raise RuntimeError(f"unexpected set_random caller chain {up}")

    def random(self):
        f = _caller(1)
        name = f.f_code.co_name
        if name == "set_random":                    # grid.py:77
            slot = CTX.pending_food_slot
            CTX.pending_food_slot = None
            if slot is None:
                return 0.0                          # p == 1 call sites: always accepted
            return uniform(CTX.bits("FOOD_ACCEPT", slot))
        if name == "_init_food":                    # environment.py:760
            entity = f.f_locals["entity"].__name__
            site = "RESET_FOOD_TRIAL" if entity == "Food" else "RESET_POISON_TRIAL"
            return uniform(CTX.bits(site, f.f_locals["i"]))
        raise RuntimeError(f"unexpected np.random.random() caller {name}")


class _NpProxy:
    """numpy with `.random` swapped for the counter generator."""

    def __init__(self):
        self.random = _NpRandom()

    def __getattr__(self, name):
        return getattr(_np, name)


class _BrainRandom:
    """`random` inside the DQN-family brain modules during the act phase; index = agent slot."""

    def random(self):
        return uniform(CTX.bits("ACT_EXPLORE", CTX.slot))

    def choice(self, seq):                          # PERD3QN.py:209, D3QN.py:172
        seq = list(seq)
        return seq[below(CTX.bits("ACT_RANDOM", CTX.slot), len(seq))]

    def randint(self, a, b):                        # DQN.py:137 (inclusive bounds)
        return a + below(CTX.bits("ACT_RANDOM", CTX.slot), b - a + 1)


class CounterRandom(__import__("random").Random):
    """CPython's own Random (so .sample / .choice run the stdlib algorithms unchanged) whose _randbelow is served by
    a counter stream: the c-th call returns below(bits_fn(c), n).  Used to drive random.sample of D3QN.py:140 /
    DQN.py:100 with the draws of RL_SITE_REPLAY_SAMPLE_UNIFORM."""

    def __init__(self, bits_fn):
        super().__init__(0)
        self._bits_fn, self.calls = bits_fn, 0

    def _randbelow(self, n):
        v = below(self._bits_fn(self.calls), n)
        self.calls += 1
        return v


def uniform_sample_bits(key, t, event_rank, it=0, n_iter=1):
    """bits_fn of one (event, iteration) stream of RL_SITE_REPLAY_SAMPLE_UNIFORM (include/rl_rng.h)."""
    base = ((event_rank * n_iter + it) << 9)
    return lambda c: draw(key, t, SITE["REPLAY_SAMPLE_UNIFORM"], base + c)


_LOADED = {}


def load_reference():
    """Import the reference with stubs + shims installed.  Returns the ReinLife package."""
    if "pkg" in _LOADED:
        return _LOADED["pkg"]
    if not os.path.isdir(os.path.join(REF_PATH, "ReinLife")):
        raise FileNotFoundError(f"reference not found at {REF_PATH}")
    for name in ("pygame", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.pyplot"].Figure = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_PATH not in sys.path:
        sys.path.insert(0, REF_PATH)
    import ReinLife  # noqa
    env_mod = sys.modules["ReinLife.World.environment"]
    grid_mod = sys.modules["ReinLife.World.grid"]
    env_mod.random = _EnvRandom()
    env_mod.np = _NpProxy()
    grid_mod.np = _NpProxy()
    _LOADED["pkg"] = ReinLife
    return ReinLife


def shim_brain_rng(enable=True):
    """Route the brains' python-`random` draws through the counter generator (act phase)."""
    load_reference()
    import random as _random
    for m in ("ReinLife.Models.PERD3QN", "ReinLife.Models.D3QN", "ReinLife.Models.DQN"):
        sys.modules[m].random = _BrainRandom() if enable else _random


# ---------------------------------------------------------------- canonical SoA dump
F_KILLED, F_INTER, F_INTRA, F_ATE, F_REPRODUCED, F_DEAD = 1, 2, 4, 8, 16, 32
REC_DTYPE = _np.dtype([("cell", "<u2"), ("health", "<i2"), ("age", "<i2"), ("max_age", "<i2"),
                       ("gene", "<i4"), ("flags", "u1"), ("action", "i1"), ("prev_slot", "<u2")])
assert REC_DTYPE.itemsize == 16


def agent_flags(a):
    return ((F_KILLED if a.killed else 0) | (F_INTER if a.inter_killed else 0) |
            (F_INTRA if a.intra_killed else 0) | (F_ATE if a.ate_super_food == 1.0 else 0) |
            (F_REPRODUCED if a.reproduced else 0) | (F_DEAD if a.dead else 0))


def dump_env(env, with_obs=True):
    """Object grid -> dict(type u8[H,W], rec[n] (row-major), reward f64[n], obs f64[n,153], state f64[n,153])."""
    H, W = env.height, env.width
    typ = _np.zeros((H, W), _np.uint8)
    recs, rewards, obs, state = [], [], [], []
    for i in range(H):
        for j in range(W):
            e = env.grid.grid[i, j]
            typ[i, j] = int(e.entity_type)
            if int(e.entity_type) == 3:
                assert (e.i, e.j) == (i, j)
                recs.append((i * W + j, e.health, e.age, e.max_age, e.gene, agent_flags(e),
                             e.action, getattr(e, "_slot_a", 0xFFFF)))
                rewards.append(float(e.reward) if e.reward is not None else 0.0)
                if with_obs:
                    obs.append(_np.asarray(e.state_prime, _np.float64))
                    state.append(_np.asarray(e.state, _np.float64))
    out = dict(type=typ, rec=_np.array(recs, dtype=REC_DTYPE), reward=_np.array(rewards, _np.float64))
    if with_obs:
        out["obs"] = _np.array(obs, _np.float64).reshape(len(recs), 153)
        out["state"] = _np.array(state, _np.float64).reshape(len(recs), 153)
    return out


class RefWorld:
    """One reference Environment = one world, driven phase by phase with teacher-forced actions."""

    def __init__(self, brains, seed=0, world=0, width=30, height=30, max_agents=100, static_families=True,
                 limit_reproduction=False, incentivize_killing=True, training=False):
        pkg = load_reference()
        self.pkg = pkg
        self.seed, self.world = seed, world
        self.env = pkg.Environment(width=width, height=height, brains=brains, grid_size=24, max_agents=max_agents,
                                   update_interval=10 ** 9, print_results=False, static_families=static_families,
                                   training=training, limit_reproduction=limit_reproduction,
                                   incentivize_killing=incentivize_killing)
        self.t = 0

    def _enter(self, t):
        CTX.set_world(self.seed, self.world)
        CTX.begin(t)

    def _mark_slots(self):
        for s, a in enumerate(self.env.agents):
            a._slot_a = s
        self._number_new_agents()

    def _number_new_agents(self):
        """Object identity for `agent not in self.best_agents` (environment.py:738): agents that are new at the end of
        reset / update_env / top-up get consecutive serial numbers in row-major order."""
        for a in self.env.grid.get_entities(self.env.entities.agent):
            if not hasattr(a, "_serial"):
                a._serial = self._next_serial
                self._next_serial += 1

    def ns_state(self):
        """(fitness f64[n], serial i64[n]) of the listed agents + (best serial, best fitness, best brain id)[10] + max_gene
        (non-static families).  Brain ids: the gene of the lineage, or -1-k for the private copy of initial best agent k."""
        env = self.env
        agents = env.grid.get_entities(env.entities.agent)
        fit = _np.array([float(a.fitness) for a in agents], _np.float64)
        ser = _np.array([a._serial for a in agents], _np.int64)
        best = [(b._serial, float(b.fitness), b.gene if b._serial >= 0 else b._serial) for b in env.best_agents]
        return dict(fitness=fit, serial=ser, best_serial=_np.array([b[0] for b in best], _np.int64),
                    best_fitness=_np.array([b[1] for b in best], _np.float64),
                    best_brain=_np.array([b[2] for b in best], _np.int32), max_gene=int(env.max_gene))

    def reset(self):
        self._enter(0)
        self._next_serial = 0
        self.env.reset()
        for k, b in enumerate(self.env.best_agents):        # the ten deep copies of agent 0 (environment.py:149)
            b._serial = -1 - k
        self._mark_slots()
        self.t = 0

    def load_state(self, typ, rec):
        """Build an arbitrary world directly (crafted scenarios).  No observation is computed."""
        from ReinLife.World.grid import Grid
        from ReinLife.World.entities import Agent, Food, Poison, SuperFood
        env = self.env
        H, W = env.height, env.width
        env.grid = Grid(W, H)
        cls = {1: Food, 2: Poison, 5: SuperFood}
        for i in range(H):
            for j in range(W):
                tt = int(typ[i, j])
                if tt in cls:
                    env.grid.set(i, j, cls[tt])
        for r in rec:
            i, j = divmod(int(r["cell"]), W)
            g = int(r["gene"])
            a = env.grid.set(i, j, Agent, brain=env.brains[g] if g < len(env.brains) else None, gene=g)
            a.health, a.age, a.max_age = int(r["health"]), int(r["age"]), int(r["max_age"])
            f = int(r["flags"])
            a.killed, a.inter_killed, a.intra_killed = int(bool(f & 1)), int(bool(f & 2)), int(bool(f & 4))
            a.ate_super_food = 1.0 if f & 8 else -1
            a.reproduced = bool(f & 16)
            a.dead = bool(f & 32)
            a.action = int(r["action"])
        env.agents = env.grid.get_entities(env.entities.agent)
        if not hasattr(self, "_next_serial"):
            self._next_serial = 0
        self._mark_slots()

    def observe(self):
        self.env._get_observations()
        self.env._update_agents_state()
        self._mark_slots()

    def force_actions(self, actions):
        assert len(actions) == len(self.env.agents)
        for a, agent in zip(actions, self.env.agents):
            agent.action = int(a)

    def step(self):
        self.t += 1
        self._enter(self.t)
        self.env.step()

    def update_env(self, n_epi=0):
        self._enter(self.t)
        self.env.update_env(n_epi)
        self._mark_slots()

    def top_up(self, target, max_age=50):
        """SURVEY 8d saturated generator, keyed by the TOPUP_* sites; then re-observe."""
        from ReinLife.World.entities import Agent
        env = self.env
        self._enter(self.t)
        n = len(env.grid.get_entities(env.entities.agent))
        k = 0
        while n < target:
            grid = env.grid.get_numpy()
            ii, jj = _np.where(grid == 0)
            if len(ii) == 0:
                break
            e = below(CTX.bits("TOPUP_PLACE", k), len(ii))
            g = below(CTX.bits("TOPUP_GENE", k), len(env.brains))
            a = env.grid.set(int(ii[e]), int(jj[e]), Agent, brain=env.brains[g], gene=g)
            a.health = 10 * (1 + below(CTX.bits("TOPUP_HEALTH", k), 20))
            a.age = below(CTX.bits("TOPUP_AGE", k), max_age)
            n += 1
            k += 1
        env._get_observations()
        env._update_agents_state()
        self._mark_slots()

    def dump(self, **kw):
        return dump_env(self.env, **kw)


class NullBrain:
    """Minimal BasicBrain stand-in for world-only goldens (never asked to act or learn)."""
    method = "PERD3QN"
    input_dim, output_dim = 153, 8

    def apply_gaussian_noise(self):                 # Agent.mutate_brain (entities.py:210-213) after a non-static _produce
        pass

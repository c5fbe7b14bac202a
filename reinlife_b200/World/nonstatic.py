"""Environment(static_families=False) -- the reference's evolving-lineage mode (World/environment.py:149, 506-507,
541-547, 728-739) on the device World, with a brain pool per world.

What runs where:
* World (reset / step / update_env incl. Agent.fitness, object identity, the ten best agents, `_update_best_agents`, the
  non-static `_produce`) = the batched CUDA kernels `k_world_*<NS>` over all worlds at once (csrc/world_kernels.cu,
  bit-exact against the reference: tests/test_world_ns_gpu.py).  update_env leaves the `_produce` event of every world --
  (new gene, source brain id) -- in `rl_ns_state`.
* Brains: a brain is a per-lineage OBJECT created inside one world and never shared across worlds ("replicas
  only"): offspring share the parent's brain (:506-507), a produced agent gets `deepcopy(random.choice(best_agents).brain)`
  -- weights, target, optimizer state, replay memory, epsilon -- followed by `mutate_brain()` (:543-547).  Here every
  lineage brain is a reinlife_b200.Models brain in plugin mode (reinlife_b200/plugin.py: private device ring, networks and
  Adam state; get_action / learn with the reference's per-agent order of effects, pinned in tests/test_seq_gpu.py), deep
  copies are device-to-device tensor clones, and the act / learn phases are the reference's own per-agent loops
  (Helpers/trainer.py:88-96) driven from one host snapshot of the device World per phase.
  This is the exact-semantics path: cost is per agent (a handful of small launches), so it serves the reference's use
  (one or a few worlds); the batched per-slot brain kernels for thousands of non-static worlds are not built.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .vecworld import VecWorld
from .utils import Actions, EntityTypes

VARIABLES = ["Avg Population Size", "Avg Population Age", "Avg Population Fitness", "Best Population Age",
             "Avg Number of Attacks", "Avg Number of Kills", "Avg Number of Intra Kills", "Avg Number of Populations"]


class _BestAgent:
    """What Saver.save reads from a best agent (environment.py:252-256): `.brain`, `.gene`."""

    def __init__(self, gene, brain, fitness):
        self.gene, self.brain, self.fitness = gene, brain, fitness


class _NSTracker:
    """Helpers/tracker.py with static_families=False: one pooled 'gene' (nr_genes = 1), every series over ALL agents of a
    world (:178-266); with N worlds the per-step value is the mean over the worlds that have agents."""

    def __init__(self, update_interval, print_results):
        self.update_interval, self.print_results = int(update_interval), print_results
        self.nr_genes = 1
        self.results = {v: ({0: []} if v != VARIABLES[-1] else []) for v in VARIABLES}
        self.track = {v: [] for v in VARIABLES}
        self.variables = list(self.results.keys())

    @staticmethod
    def world_series(rec, reward):
        if len(rec) == 0:
            return [-1] * 8
        _, counts = np.unique(rec["gene"], return_counts=True)
        killed = int((rec["flags"] & _lib.F_KILLED != 0).sum())
        return [float(np.mean(counts)), float(np.mean(rec["age"])), float(np.mean(reward.astype(np.float64))), int(rec["age"].max()),
                float((rec["action"] >= 4).sum()) / len(rec), killed, (1.0 if killed else 0), len(counts)]

    def update_results(self, per_world, n_epi):
        vals = [s for s in per_world if s[0] != -1]
        row = [float(np.mean([s[i] for s in vals])) for i in range(8)] if vals else [-1] * 8
        for v, x in zip(VARIABLES, row):
            self.track[v].append(x)
        if n_epi % self.update_interval == 0 and n_epi != 0:
            for v in VARIABLES:
                xs = [x for x in self.track[v][-self.update_interval:] if x > -1]
                agg = float(np.mean(xs)) if xs else float("nan")
                (self.results[v][0] if v != VARIABLES[-1] else self.results[v]).append(agg)
                self.track[v] = []
            if self.print_results:
                print("################ non-static families ################")
                for v in VARIABLES:
                    r = self.results[v][0][-1] if v != VARIABLES[-1] else self.results[v][-1]
                    print(f"{v:28s}  {r:8.3f}")


class NonStaticEnvironment:
    def __init__(self, width=30, height=30, brains=None, grid_size=16, max_agents=50, update_interval=500, print_results=True,
                 static_families=False, interactive_results=False, google_colab=False, training=True, save=False,
                 pastel_colors=False, limit_reproduction=False, incentivize_killing=True, *, n_worlds=1, seed=0, device=None,
                 world_id0=None, precision="fp32", sequential_events=True, slot_cap=None):
        self.width, self.height = width, height
        self.actions, self.entities = Actions, EntityTypes
        self.brains = brains
        self.max_agents = max_agents
        self.max_gene = len(brains)                      # TypeError when brains is None, like environment.py:107
        self.static_families, self.google_colab, self.save = False, google_colab, save
        self.training, self.limit_reproduction, self.incentivize_killing = training, limit_reproduction, incentivize_killing
        self.action_space, self.observation_space = 8, 153
        self.update_interval, self.print_results = update_interval, print_results
        if interactive_results:
            raise NotImplementedError("interactive matplotlib results are out of scope (SURVEY.md 2, #16)")
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            raise NotImplementedError("static_families=False: replicas only -- run one process per GPU, no sharding (SURVEY.md 8e)")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RLError("reinlife_b200 runs on CUDA devices only (no CPU fallback)")
        for g, b in enumerate(brains):
            if not hasattr(b, "_plugin_learn"):
                raise TypeError(f"brain {g} ({type(b).__name__}) is not a reinlife_b200.Models brain")
            if getattr(b, "_env", None) is not None and b._plugin_host is None:
                raise RuntimeError("a brain bound to a vectorised Environment cannot seed a non-static one")
        self.n_worlds = self.n_worlds_global = int(n_worlds)
        self.seed = seed
        self.world = VecWorld(self.n_worlds, height, width, len(brains), max_agents=max_agents, seed=seed,
                              world_id0=world_id0 or 0, static_families=False, limit_reproduction=limit_reproduction,
                              incentivize_killing=incentivize_killing, device=self.device, slot_cap=slot_cap)
        self.tracker = _NSTracker(update_interval, print_results)
        self.pools = None           # per world: {brain id: brain}; ids = lineage genes, -1-k = private copy of initial best agent k
        self.viz = None
        self.grid = None
        self.gpu_launches = 0
        self.precision = "fp32"
        self._state = None          # host snapshot of obs_state at the act phase (agent.state of the step's transitions)
        self._probs = None          # PPO: pi(a) tensors of the act phase, per (world, slot)

    # ------------------------------------------------------------------ reference phases
    def reset(self):
        self.world.reset()
        G = len(self.brains)
        # world 0 owns the caller's brain objects (like the reference); every further world is an independent run with its
        # own deep copies.  best_agents = ten deep copies of agent 0 (:149): ONE snapshot of brain 0 serves the ten ids.
        self.pools = []
        for w in range(self.n_worlds):
            pool = {g: (self.brains[g] if w == 0 else self.brains[g].clone()) for g in range(G)}
            pool["snapshot"] = pool[0].clone()
            self.pools.append(pool)
        self.max_gene = G

    def _brain(self, w, brain_id):
        pool = self.pools[w]
        return pool["snapshot"] if brain_id < 0 else pool[brain_id]

    def act(self, n_epi=0):
        """for agent in env.agents: agent.get_action(n_epi)   (Helpers/trainer.py:88-89, World/entities.py:215-222)"""
        w = self.world
        n, rec = self.snapshot_state()
        acts = np.zeros((self.n_worlds, w.S), np.int8)
        for wi in range(self.n_worlds):
            for s in range(int(n[wi])):
                brain = self.pools[wi][int(rec[wi, s]["gene"])]
                state = self._state[wi, s, :_lib.OBS_DIM].astype(np.float64)
                if brain.method == "PPO":
                    out = brain.get_action(state)
                    a, prob = out if isinstance(out, tuple) else (out, None)
                    self._probs[(wi, s)] = prob
                elif brain.method == "PERDQN":
                    a = brain.get_action(state)
                else:
                    a = brain.get_action(state, n_epi)
                acts[wi, s] = a
        w.set_actions(acts)

    def snapshot_state(self):
        """Host copy of `agent.state` of every listed agent (the `state` of the transitions this step will produce)."""
        w = self.world
        torch.cuda.synchronize(self.device)
        self._state = w.obs_state.cpu().numpy()
        self._probs = {}
        return w.n_agents.cpu().numpy(), w.rec_host()

    def step(self):
        self.world.step()

    def learn(self, n_epi=0):
        """for agent in env.agents: agent.learn(n_epi=n_epi)   (Helpers/trainer.py:95-96, World/entities.py:194-208)"""
        if not self.training:
            return
        w = self.world
        torch.cuda.synchronize(self.device)
        n = w.n_agents.cpu().numpy()
        rec = w.rec_host()
        prime = w.obs_prime.cpu().numpy()
        reward = w.reward.cpu().numpy()
        for wi in range(self.n_worlds):
            for s in range(int(n[wi])):
                r = rec[wi, s]
                if int(r["age"]) <= 1:                               # World/entities.py:196
                    continue
                brain = self.pools[wi][int(r["gene"])]
                dead = bool(r["flags"] & _lib.F_DEAD)
                prev = int(r["prev_slot"])
                kw = dict(age=int(r["age"]), dead=dead, action=int(r["action"]), state=self._state[wi, prev, :_lib.OBS_DIM],
                          reward=float(reward[wi, s]), state_prime=prime[wi, s, :_lib.OBS_DIM], done=dead)
                if brain.method == "PPO":
                    brain.learn(prob=self._probs[(wi, prev)], **kw)
                elif brain.method in ("DQN", "PERDQN"):
                    brain.learn(**kw)
                else:
                    brain.learn(n_epi=n_epi, **kw)

    def update_env(self, n_epi=0):
        w = self.world
        if self.training:                                            # environment.py:206-207
            torch.cuda.synchronize(self.device)
            n = w.n_agents.cpu().numpy(); rec = w.rec_host(); rew = w.reward.cpu().numpy()
            self.tracker.update_results([_NSTracker.world_series(rec[wi, :n[wi]], rew[wi, :n[wi]]) for wi in range(self.n_worlds)], n_epi)
        w.update()
        torch.cuda.synchronize(self.device)
        states = w.ns_host()
        n = w.n_agents.cpu().numpy()
        rec = w.rec_host()
        for wi in range(self.n_worlds):
            st, pool = states[wi], self.pools[wi]
            if st.produced_gene >= 0:                                # _produce (:541-547): deep copy of a random best agent's brain
                new = self._brain(wi, st.produced_src_brain).clone()
                placed = bool((rec[wi, :n[wi]]["gene"] == st.produced_gene).any())
                if placed and new.method == "PERD3QN":               # agent.mutate_brain() (World/entities.py:210-213)
                    new.apply_gaussian_noise()
                pool[st.produced_gene] = new
            live = set(int(g) for g in rec[wi, :n[wi]]["gene"]) | set(b.brain for b in st.best if b.brain >= 0)
            for g in [g for g in pool if g != "snapshot" and g not in live]:
                del pool[g]                                          # no agent and no best-table entry refers to this brain any more
        self.max_gene = max(int(s.max_gene) for s in states)

    def render(self, fps=10):
        raise NotImplementedError("the pygame renderer is out of scope (SURVEY.md 2, #18)")

    # ------------------------------------------------------------------ host views
    @property
    def best_agents(self):
        """World 0's ten best agents (environment.py:728-739) as objects with `.brain`, `.gene`, `.fitness`."""
        st = self.world.ns_host()[0]
        return [_BestAgent(b.brain if b.brain >= 0 else 0, self._brain(0, b.brain), b.fitness) for b in st.best]

    def save_results(self):
        """environment.py:233-256 with families=False: the best agents' brains as brain_1, brain_2, ... per method."""
        from ..Helpers.saver import Saver
        settings = {"Update interval": self.update_interval, "Width": self.width, "Height": self.height,
                    "Max agents": self.max_agents, "Families": False}
        return Saver("experiments", google_colab=self.google_colab).save(self.best_agents, False, self.tracker.results, settings, None)

    def check_status(self):
        st = int(self.world.status.max())
        if st & 1:
            raise RuntimeError("a world overflowed slot_cap")
        if st & 4:
            raise RuntimeError("more than 256 distinct lineages alive in one world")

    def count_agents(self):
        return int(self.world.n_agents.sum())

    def lineages(self, world=0):
        return sorted(g for g in self.pools[world] if g != "snapshot")

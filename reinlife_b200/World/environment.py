"""Vectorised Environment: `n_worlds` independent instances of the reference's Environment
(ReinLife/World/environment.py:16-215) stepped together on one GPU, the brains' networks shared by all worlds.

Same constructor keywords and phase methods as the reference (`reset`, `step`, `update_env`, `render`,
`save_results`) plus the two batched phases that replace the per-agent Python loops of Helpers/trainer.py:88-96:
`act(n_epi)` and `learn(n_epi)`.  Extra keyword-only arguments: n_worlds, seed, device, world_id0.
With torch.distributed initialised, `n_worlds` is the GLOBAL world count: each rank owns a contiguous shard and the
only collective is one all-reduce of the brains' summed gradients per learn step.
"""
import ctypes as C
import os

import numpy as np
import torch

from .. import _lib
from ..brains import sync_target
from ..rows import RowLists
from .vecworld import VecWorld
from .utils import Actions, EntityTypes


def _nvtx(name):
    """Decorator: an NVTX range around a phase when RL_NVTX is set (nsys / ncu --nvtx timelines)."""
    def wrap(fn):
        if os.environ.get("RL_NVTX") is None:
            return fn

        def inner(*a, **kw):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **kw)
            finally:
                torch.cuda.nvtx.range_pop()
        inner.__doc__ = fn.__doc__
        return inner
    return wrap


class AgentView:
    """Read-only host snapshot of one agent (the fields of World/entities.py:145-170 that live on the device)."""
    __slots__ = ("i", "j", "health", "max_health", "age", "max_age", "gene", "action", "killed", "inter_killed",
                 "intra_killed", "ate_super_food", "reproduced", "dead", "reward", "brain", "coordinates")

    def __init__(self, rec, width, reward, brain):
        self.i, self.j = divmod(int(rec["cell"]), width)
        self.coordinates = [self.i, self.j]
        self.health, self.max_health = int(rec["health"]), 200
        self.age, self.max_age, self.gene, self.action = int(rec["age"]), int(rec["max_age"]), int(rec["gene"]), int(rec["action"])
        f = int(rec["flags"])
        self.killed, self.inter_killed, self.intra_killed = f & 1, (f >> 1) & 1, (f >> 2) & 1
        self.ate_super_food = 1.0 if f & 8 else -1
        self.reproduced, self.dead = bool(f & 16), bool(f & 32)
        self.reward, self.brain = reward, brain


class _GridView:
    def __init__(self, env, world=0):
        self._env, self._world = env, world
        self.width, self.height = env.width, env.height

    def get_numpy(self, entity_type=None):
        g = self._env.world.type[self._world].cpu().numpy().reshape(self.height, self.width).astype(np.int64)
        return (g == entity_type) if entity_type else g


class Environment:
    def __new__(cls, *args, **kwargs):
        # static_families=False (8th positional argument of the reference's constructor, environment.py:74-89) is the
        # evolving-lineage mode: reinlife_b200.World.nonstatic.NonStaticEnvironment (device World + per-world brain pools)
        static = kwargs.get("static_families", args[7] if len(args) > 7 else True)
        if cls is Environment and not static:
            from .nonstatic import NonStaticEnvironment
            return NonStaticEnvironment(*args, **kwargs)
        return super().__new__(cls)

    def __init__(self, width: int = 30, height: int = 30, brains=None, grid_size: int = 16, max_agents: int = 50,
                 update_interval: int = 500, print_results: bool = True, static_families: bool = True,
                 interactive_results: bool = False, google_colab: bool = False, training: bool = True,
                 save: bool = False, pastel_colors: bool = False, limit_reproduction: bool = False,
                 incentivize_killing: bool = True, *, n_worlds: int = 1, seed: int = 0, device=None, world_id0=None,
                 precision: str = "fp16", sequential_events: bool = False):
        self.width, self.height = width, height
        self.actions, self.entities = Actions, EntityTypes
        self.best_agents = []
        self.brains = brains
        self.max_agents = max_agents
        self.max_gene = len(brains)                      # TypeError when brains is None, like environment.py:107
        self.static_families, self.google_colab, self.save = static_families, google_colab, save
        self.training, self.limit_reproduction, self.incentivize_killing = training, limit_reproduction, incentivize_killing
        self.action_space, self.observation_space = 8, 153
        self.update_interval, self.print_results = update_interval, print_results
        if interactive_results:
            raise NotImplementedError("interactive matplotlib results are out of scope (SURVEY.md 2, #16)")

        self.dist = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.rank = torch.distributed.get_rank() if self.dist else 0
        self.world_size = torch.distributed.get_world_size() if self.dist else 1
        if device is None:
            device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", torch.cuda.current_device() if torch.cuda.is_available() else 0)))
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RLError("reinlife_b200 runs on CUDA devices only (no CPU fallback)")
        from ..sharding import shard_worlds
        self.n_worlds_global = n_worlds
        self.n_worlds, wid0 = shard_worlds(n_worlds, self.rank, self.world_size)
        self.seed = seed
        wid0 = wid0 if world_id0 is None else world_id0
        self.world = VecWorld(self.n_worlds, height, width, len(brains), max_agents=max_agents, seed=seed, world_id0=wid0,
                              static_families=static_families, limit_reproduction=limit_reproduction,
                              incentivize_killing=incentivize_killing, device=self.device)
        self.rows = RowLists(self.world)
        G = len(brains)
        # "tf32": train() events on the tensor cores (tcgen05 kind::tf32, fp32 accumulate); "fp32": CUDA-core FMA path
        if precision not in ("tf32", "fp16", "fp32"):
            raise ValueError("precision must be 'tf32', 'fp16' or 'fp32'")
        # "fp16": train() events with fp16 operands (tcgen05 kind::f16, fp32 accumulate); get_action stays on the tf32 forward
        self._learn_fp16 = precision == "fp16"
        self.precision = "tf32" if precision == "fp16" else precision
        self._learn_single = os.environ.get("RL_LEARN_SINGLE") is not None   # A/B: the one-event-per-iteration fp16 kernel
        for g, b in enumerate(brains):
            if not hasattr(b, "_bind"):
                raise TypeError(f"brain {g} ({type(b).__name__}) is not a reinlife_b200.Models brain")
            b._bind(self, g)
        for b in brains:
            if getattr(b, "_dev", None) is not None:
                b._dev.use_fp16 = self._learn_fp16
        # fp16 events of the dueling brains: the World kernels also emit float16 rows, which get_action gathers by TMA
        # (k_act_dueling_p<true>) and rl_replay_store copies into the float16 rings (RL_NO_OBS16: A/B switch)
        if (self._learn_fp16 and not self._learn_single and os.environ.get("RL_NO_OBS16") is None
                and any(getattr(b, "KIND", None) == _lib.MODEL_DUELING for b in brains)):
            self.world.enable_obs_fp16()
        self._eps = torch.tensor([float(b.epsilon) if hasattr(b, "epsilon") else 0.0 for b in brains],
                                 dtype=torch.float64, device=self.device)
        self._seen = torch.tensor([int(getattr(b, "n_epi", 0)) for b in brains], dtype=torch.int64, device=self.device)
        self._prob = torch.zeros(self.n_worlds * self.world.S, device=self.device)
        self._sample_status = torch.zeros(1, dtype=torch.int32, device=self.device)   # bit 0: random.sample on a short buffer
        self._act_descs = (_lib.BrainAct * G)(*[b._dev.act_desc(b.RULE, self._eps.data_ptr() + 8 * g) for g, b in enumerate(brains)])
        if self.dist:
            for b in brains:                              # identical weights on every rank
                torch.distributed.broadcast(b._dev.params, 0)
                if b._dev.target is not None:
                    torch.distributed.broadcast(b._dev.target, 0)
        # sequential_events=True (one world, one rank): learn() walks the agents in the reference's order and runs
        # store -> train -> priorities -> Adam -> target sync PER AGENT, so that every train() sees the weights and the ring
        # the previous agent's train() left (Helpers/trainer.py:95-96, PERD3QN.py:117-125) -- the reference's exact N = 1
        # semantics; the default batches all events of a step against the pre-step weights (learn-step semantics)
        self.sequential_events = bool(sequential_events)
        if self.sequential_events and (self.n_worlds != 1 or self.world_size != 1):
            raise ValueError("sequential_events=True is the exact single-world mode: n_worlds must be 1 (one rank)")
        self._act_tc = os.environ.get("RL_ACT_FP32") is None    # tf32 runs: get_action of dueling brains on the tensor cores too
        self._grad_all = None
        # Multi-GPU: the "did this brain act / store / trigger this step" conditions of the epsilon schedules and target
        # syncs must be the same on every rank (the reference has ONE brain object), so they are taken from the row totals
        # summed over all ranks (one 4*3G-byte all-reduce after each row-list build), not from this rank's lists.
        self._gate = self.rows.total
        self._rows_gate = self.rows.bufs
        # row lists / gate / sampler key the learn pipeline currently works on (the per-agent pipeline of
        # sequential_events swaps in one-row views of the lists)
        self._rb_store = self._rb_event = self.rows.bufs
        self._gate_base = None
        self._t_key = None
        self.sample_override = None     # tests: callable(gene, k_event) -> 64 ring positions used instead of the sampler
        self.event_hook = None          # tests: callable(gene, k_event) after every sequential train() event
        self._seq_event = None
        if self.dist:
            self._gate = torch.zeros_like(self.rows.total)
            rb = self.rows.bufs
            self._rows_gate = _lib.RowsBufs(rb.count, rb.offset, self._gate.data_ptr(), rb.rows, rb.row_cap, 0)
        from ..Helpers.tracker import Tracker
        self.tracker = Tracker(self, update_interval=update_interval, print_results=print_results)
        self.viz = None
        self.grid = None
        self.gpu_launches = 0
        self.kernel_events = None       # set to a list to collect (name, gene, start, end) CUDA events of the hot kernels

    # ------------------------------------------------------------------ reference phases
    @_nvtx("reinlife.reset")
    def reset(self):
        self.world.reset()
        self.grid = _GridView(self)
        self.gpu_launches += 1

    @_nvtx("reinlife.step")
    def step(self):
        self.world.step()
        self.gpu_launches += 1

    @_nvtx("reinlife.update_env")
    def update_env(self, n_epi: int = 0, top_up=None, max_age=50):
        """environment.py:188-215.  top_up=N (benchmark loops, static families): the saturated-world generator runs in the same
        launch -- identical to update_env() followed by top_up(N)."""
        if self.training:                                  # environment.py:206-207
            self.tracker.update_results(None, n_epi)
            self.gpu_launches += 1
        ev0 = None
        if self.kernel_events is not None:                 # bench.py: CUDA-event time of the update (+ top-up) kernel alone (roofline)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        self.world.update(top_up=top_up, max_age=max_age)
        if ev0 is not None:
            ev1.record()
            self.kernel_events.append(("world_update", -1, ev0, ev1))
        self.gpu_launches += 1

    @_nvtx("reinlife.top_up")
    def top_up(self, target, max_age=50):
        """Benchmark-only saturated-world generator."""
        self.world.top_up(target, max_age)
        self.gpu_launches += 1

    def render(self, fps: int = 10) -> bool:
        raise NotImplementedError("the pygame renderer is out of scope (SURVEY.md 2, #18)")

    def save_results(self):
        from ..Helpers.saver import save_brains
        return save_brains(self)

    # ------------------------------------------------------------------ batched stand-ins for the per-agent loops
    @_nvtx("reinlife.act")
    def act(self, n_epi: int = 0, q_out=None):
        """for agent in env.agents: agent.get_action(n_epi)   (Helpers/trainer.py:88-89, Helpers/tester.py:58-68)"""
        w = self.world
        G = len(self.brains)
        self.rows.build(kinds_mask=1)
        self._reduce_gate()
        sched = (_lib.BrainSched * G)(*[b._sched() for b in self.brains])
        with torch.cuda.device(self.device):
            _lib.check(w.lib.rl_brain_epsilon_update(C.byref(self._rows_gate), sched, G, C.c_int64(n_epi),
                                                     C.c_void_p(self._eps.data_ptr()), C.c_void_p(self._seen.data_ptr()),
                                                     w._stream()))
            tc = [g for g, b in enumerate(self.brains) if self.precision == "tf32" and b.KIND == _lib.MODEL_DUELING and self._act_tc]
            # DQN-layout brains (DQN, PERDQN) under precision="fp16": k_act_dqn_p (q_out is not produced on this path)
            tcq = [g for g, d in enumerate(self._act_descs) if self._learn_fp16 and d.kind == _lib.MODEL_DQN and self._act_tc and q_out is None]
            descs = self._act_descs
            if tc or tcq:                            # these brains act on the tensor cores; the others on the fp32 path
                descs = (_lib.BrainAct * G)(*[_lib.BrainAct(-1 if (g in tc or g in tcq) else d.kind, d.rule, d.params, d.epsilon)
                                              for g, d in enumerate(self._act_descs)])
            for g in tcq:
                _lib.check(w.lib.rl_brain_act_dqn_p(C.byref(w.cfg), C.byref(w.bufs), C.byref(self.rows.bufs), C.c_int32(g),
                                                    C.byref(self._act_descs[g]), C.c_uint64(w.t + 1), None, w._stream()))
            if len(tc) + len(tcq) < G:
                _lib.check(w.lib.rl_brain_act_all(C.byref(w.cfg), C.byref(w.bufs), C.byref(self.rows.bufs), descs, G,
                                                  C.c_uint64(w.t + 1), C.c_void_p(q_out), C.c_void_p(self._prob.data_ptr()), w._stream()))
            for g in tc:
                b = self.brains[g]
                if b._dev.wimg_stale:
                    b._dev.build_wimg(w._stream())
                    b._dev.wimg_stale = False
                    self.gpu_launches += 2
                if self._learn_fp16:                     # fp16 operands: batch-major 128-row tiles (k_act_dueling_p); RL_LEARN_SINGLE: k_act_dueling_h
                    fn = w.lib.rl_brain_act_h if self._learn_single else w.lib.rl_brain_act_p
                    _lib.check(fn(C.byref(w.cfg), C.byref(w.bufs), C.byref(self.rows.bufs), C.c_int32(g),
                                                    C.byref(self._act_descs[g]), C.c_void_p(b._dev.wimg_eh.data_ptr()),
                                                    C.c_uint64(w.t + 1), C.c_void_p(q_out), w._stream()))
                else:
                    _lib.check(w.lib.rl_brain_act_tc(C.byref(w.cfg), C.byref(w.bufs), C.byref(self.rows.bufs), C.c_int32(g),
                                                     C.byref(self._act_descs[g]), C.c_void_p(b._dev.wimg_e.data_ptr()),
                                                     C.c_uint64(w.t + 1), C.c_void_p(q_out), w._stream()))
        self.gpu_launches += 4 + G

    @_nvtx("reinlife.learn")
    def learn(self, n_epi: int = 0):
        """for agent in env.agents: agent.learn(n_epi=n_epi)   (Helpers/trainer.py:95-96), batched:
        all stores of the step, then every train() trigger of every world as one event against the same pre-step
        weights, per-event gradients averaged (all-reduced across ranks), ONE optimizer step per brain and per
        reference optimizer step (1 for D3QN/PERD3QN, 5 for DQN's train(), k_epoch for PPO), then priorities / target
        sync exactly where the reference does them."""
        if not self.training:
            return
        w, lib = self.world, self.world.lib
        trainable = [g for g, b in enumerate(self.brains) if b._trains()]
        for g, b in enumerate(self.brains):
            if g not in trainable and getattr(b, "training", True):
                raise NotImplementedError(f"{b.method}: the learn step is not implemented on the device yet "
                                          "(inference / tester path only); there is no CPU fallback")
        if not trainable:
            return
        tf = [int(getattr(b, "train_freq", 1)) for b in self.brains]
        on = [int(b._trains() and n_epi > getattr(b, "exploration", -1)) for b in self.brains]
        self.rows.build(kinds_mask=6, train_freq=tf, event_on=on)
        self.gpu_launches += 3
        if self.sequential_events:
            return self._learn_sequential(n_epi, trainable, tf, on)
        self._reduce_gate()
        self._rb_store = self._rb_event = self.rows.bufs
        self._gate_base, self._t_key = self._gate.data_ptr(), w.t
        self._learn_lists(trainable, tf, on, n_epi)

    def _learn_lists(self, trainable, tf, on, n_epi):
        """The learn pipeline of every brain in `trainable` over the row lists self._rb_store / self._rb_event."""
        w, lib = self.world, self.world.lib
        st = w._stream()
        with torch.cuda.device(self.device):
            for g in trainable:                                   # brain.memorize / put_data for every age > 1 agent
                b = self.brains[g]
                if b.KIND == _lib.MODEL_PPO:
                    continue                                      # rl_ppo_store (append-only data list) in _learn_ppo
                if b.method == "PERDQN":                          # Memory.add (PERDQN.py:275-277) walks the ring's write pointer
                    _lib.check(lib.rl_sumtree_add(C.byref(w.cfg), C.byref(self._rb_store), C.c_int32(g), C.byref(b._replay.bufs),
                                                  C.byref(b.memory.bufs), st))
                    self.gpu_launches += 1
                _lib.check(lib.rl_replay_store(C.byref(w.cfg), C.byref(w.bufs), C.byref(self._rb_store), C.c_int32(g),
                                               C.byref(b._replay.bufs), st))
                self.gpu_launches += 1
            self._learn_dueling([g for g in trainable if on[g] and self.brains[g].KIND == _lib.MODEL_DUELING], n_epi, st)
            for g in trainable:
                if self.brains[g].method == "PERDQN":
                    self._learn_perdqn(g, st)
                elif on[g] and self.brains[g].KIND == _lib.MODEL_DQN:
                    self._learn_dqn(g, st)
                elif self.brains[g].KIND == _lib.MODEL_PPO:
                    self._learn_ppo(g, tf[g], st)

    # ------------------------------------------------------------------ exact single-world semantics
    def _one_row_views(self, k_store, k_event, trigger):
        """Row-list structs whose STORE list is [k_store-th STORE row] and whose EVENT list is [k_event-th EVENT row]
        (empty when the agent does not trigger): pointer arithmetic on the lists rl_rows_build wrote, no copies."""
        rb = self.rows.bufs
        c1, c0 = self._seq_c1.data_ptr(), self._seq_c0.data_ptr()
        self._rb_store = _lib.RowsBufs(c1, c0, c1, rb.rows + 4 * k_store, rb.row_cap, 0)
        ce = c1 if trigger else c0
        self._rb_event = _lib.RowsBufs(ce, c0, ce, rb.rows + 4 * k_event, rb.row_cap, 0)
        self._seq_gate[_lib.ROWS_STORE::_lib.N_ROW_KINDS] = 1
        self._seq_gate[_lib.ROWS_EVENT::_lib.N_ROW_KINDS] = 1 if trigger else 0
        self._gate_base = self._seq_gate.data_ptr()

    def _learn_sequential(self, n_epi, trainable, tf, on):
        """for agent in env.agents: agent.learn(n_epi=n_epi) with the reference's per-agent order of effects
        (Helpers/trainer.py:95-96 -> World/entities.py:194-208 -> brain.learn)."""
        G = len(self.brains)
        if not hasattr(self, "_seq_c1"):
            K = G * _lib.N_ROW_KINDS
            self._seq_c1 = torch.ones(K, dtype=torch.int32, device=self.device)
            self._seq_c0 = torch.zeros(K, dtype=torch.int32, device=self.device)
            self._seq_gate = torch.zeros(K, dtype=torch.int32, device=self.device)
        n = int(self.world.n_agents[0])
        rec = self.world.rec_host()[0, :n]
        k_store, k_event = [0] * G, [0] * G
        for s in range(n):                                        # row-major = the reference's agent order
            g, age, dead = int(rec["gene"][s]), int(rec["age"][s]), bool(rec["flags"][s] & _lib.F_DEAD)
            if age <= 1 or g not in trainable:                    # World/entities.py:196: learn only if age > 1
                continue
            trigger = bool(on[g]) and (age % tf[g] == 0 or dead)
            self._learn_agent(g, k_store[g], k_event[g], trigger, tf, on, n_epi)
            k_store[g] += 1
            k_event[g] += int(trigger)
        self._rb_store = self._rb_event = self.rows.bufs

    def _learn_agent(self, g, k_store, k_event, trigger, tf, on, n_epi):
        self._one_row_views(k_store, k_event, trigger)
        self._t_key = (self.world.t << 20) | k_event              # sampler draws: a stream per (step, event) -- rl_rng.h
        self._seq_event = (g, k_event) if trigger else None
        self._learn_lists([g], tf, on, n_epi)
        self._seq_event = None
        if trigger and self.event_hook is not None:
            self.event_hook(g, k_event)

    def _learn_dueling(self, active, n_epi, st):
        """PERD3QN / D3QN: learn() -> train() (PERD3QN.py:94-125, D3QN.py:97-126)."""
        if not active:
            return
        w, lib = self.world, self.world.lib
        pending = []
        for g in active:
            b = self.brains[g]
            if b.PRIORITIZED:                                   # PERD3QN.py:157-175
                _lib.check(lib.rl_replay_sample(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                                C.c_int32(b._dev.batch), C.c_uint64(self._t_key), C.c_void_p(b._dev.sample_idx.data_ptr()), st))
            else:                                               # random.sample(deque, 64), D3QN.py:138-142
                _lib.check(lib.rl_replay_sample_uniform(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                                        C.c_int32(b._dev.batch), C.c_uint64(self._t_key), C.c_int32(0), C.c_int32(1), C.c_int32(0),
                                                        C.c_void_p(b._dev.sample_idx.data_ptr()), C.c_void_p(self._sample_status.data_ptr()), st))
            if self.sequential_events and not b.PRIORITIZED and getattr(self, "_seq_event", None) is not None:
                if int(self._sample_status) & 1:                      # the reference raises inside train(), before any update
                    self._sample_status.zero_()
                    raise ValueError("Sample larger than population or is negative")     # random.sample, D3QN.py:140
            if self.sample_override is not None and getattr(self, "_seq_event", None) is not None:
                idx = self.sample_override(*self._seq_event)          # tests: the reference's recorded np.random.choice result
                b._dev.sample_idx[0].copy_(torch.as_tensor(idx, dtype=torch.int32))
            ev0 = ev1 = None
            if self.kernel_events is not None:           # bench.py: CUDA-event time of the event kernel alone (roofline)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if self.precision == "tf32":
                if b._dev.wimg_stale:
                    b._dev.build_wimg(st)
                    b._dev.wimg_stale = False
                    self.gpu_launches += 2
                if ev0 is not None:
                    ev0.record()
                if self._learn_fp16:                                  # two events per CTA iteration (csrc/tc_pair_kernels.cu)
                    fn = lib.rl_brain_learn_h if self._learn_single else lib.rl_brain_learn_p
                    _lib.check(fn(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                                    C.c_void_p(b._dev.sample_idx.data_ptr()), C.byref(b._dev.learn_bufs),
                                                    C.c_void_p(b._dev.wimg_eh.data_ptr()), C.c_void_p(b._dev.wimg_th.data_ptr()), st))
                else:
                    _lib.check(lib.rl_brain_learn_tc(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                                     C.c_void_p(b._dev.sample_idx.data_ptr()), C.byref(b._dev.learn_bufs),
                                                     C.c_void_p(b._dev.wimg_e.data_ptr()), C.c_void_p(b._dev.wimg_t.data_ptr()), st))
            else:
                if ev0 is not None:
                    ev0.record()
                _lib.check(lib.rl_brain_learn(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                              C.c_void_p(b._dev.sample_idx.data_ptr()), C.byref(b._dev.learn_bufs), st))
            if ev0 is not None:                          # brackets the event kernel + its 0.03 ms slab reduction
                ev1.record()
                self.kernel_events.append(("learn_events", g, ev0, ev1))
            self.gpu_launches += 3
            if self.dist:                                # this brain's gradient all-reduce runs under the next brain's event kernel
                pending.append(torch.distributed.all_reduce(b._dev.grad, async_op=True))
        for work in pending:
            work.wait()
        for g in active:
            b = self.brains[g]
            _lib.check(lib.rl_brain_adam(C.byref(b._dev.learn_bufs), st))
            self.gpu_launches += 2
            if b.PRIORITIZED:
                _lib.check(lib.rl_replay_update_prio(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                                     C.c_int32(b._dev.batch), C.c_void_p(b._dev.sample_idx.data_ptr()),
                                                     C.c_void_p(b._dev.new_prio.data_ptr()), st))
                self.gpu_launches += 1
            synced = n_epi % int(b.soft_update_freq) == 0
            if synced:                                        # PERD3QN.py:124-125, only if learn() was called
                cond = self._gate_base + 4 * (g * _lib.N_ROW_KINDS + _lib.ROWS_STORE)
                sync_target(b._dev, w, cond)
                self.gpu_launches += 1
            if self.precision == "tf32":                      # operand images follow the parameters
                b._dev.build_wimg(st, "both" if synced else "eval")
                self.gpu_launches += 2 if synced else 1

    def _learn_dqn(self, g, st):
        """DQN: learn() -> train() (DQN.py:78-89, 142-153): if the ring holds > 1000 items, 5 x (random.sample 32,
        smooth-L1, Adam step); then target <- agent at every trigger, trained or not."""
        w, lib, b = self.world, self.world.lib, self.brains[g]
        for it in range(5):
            _lib.check(lib.rl_replay_sample_uniform(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                                    C.c_int32(32), C.c_uint64(self._t_key), C.c_int32(it), C.c_int32(5), C.c_int32(b.min_buffer),
                                                    C.c_void_p(b._dev.sample_idx.data_ptr()), None, st))
            # precision="fp16": the tensor-core iteration kernel (csrc/tc_dqn_kernels.cu); "tf32" / "fp32": the fp32 tile kernel
            fn = lib.rl_brain_learn_dqn_p if self._learn_fp16 else lib.rl_brain_learn_dqn
            _lib.check(fn(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                              C.c_void_p(b._dev.sample_idx.data_ptr()), C.byref(b._dev.learn_bufs), st))
            if self.dist:
                self._allreduce_grads([g])
            _lib.check(lib.rl_brain_adam(C.byref(b._dev.learn_bufs), st))
            self.gpu_launches += 6
        cond = self._gate_base + 4 * (g * _lib.N_ROW_KINDS + _lib.ROWS_EVENT)
        sync_target(b._dev, w, cond)
        self.gpu_launches += 1

    def _learn_perdqn(self, g, st):
        """PERDQN: learn() (PERDQN.py:188-195): every trigger of a world whose memory holds >= train_start items is one
        train_model() event (stratified SumTree sample, importance-weighted MSE, priorities of the sampled leaves
        refreshed); epsilon steps once per optimizer step; target_model <- model at every trigger, trained or not."""
        w, lib, b = self.world, self.world.lib, self.brains[g]
        tr, dev = b.memory, b._dev
        _lib.check(lib.rl_sumtree_sample(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                         C.byref(tr.bufs), C.c_int32(64), C.c_uint64(self._t_key), C.c_void_p(dev.sample_idx.data_ptr()),
                                         C.c_void_p(tr.ev_weight.data_ptr()), st))
        fn = lib.rl_brain_learn_perdqn_p if self._learn_fp16 else lib.rl_brain_learn_perdqn
        _lib.check(fn(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                             C.c_void_p(dev.sample_idx.data_ptr()), C.c_void_p(tr.ev_weight.data_ptr()),
                                             C.byref(dev.learn_bufs), st))
        if self.dist:
            self._allreduce_grads([g])
        _lib.check(lib.rl_brain_adam(C.byref(dev.learn_bufs), st))
        _lib.check(lib.rl_sumtree_update(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                         C.byref(tr.bufs), C.c_int32(64), C.c_void_p(dev.sample_idx.data_ptr()),
                                         C.c_void_p(dev.new_prio.data_ptr()), st))
        _lib.check(lib.rl_perdqn_epsilon_step(C.byref(dev.learn_bufs), C.c_void_p(self._eps.data_ptr() + 8 * g),
                                              C.c_double(b.epsilon_min), C.c_double(b.epsilon_decay), st))
        cond = self._gate_base + 4 * (g * _lib.N_ROW_KINDS + _lib.ROWS_EVENT)
        sync_target(dev, w, cond)
        self.gpu_launches += 10

    def _learn_ppo(self, g, train_freq, st):
        """PPO: learn() (PPO.py:71-77): put_data for every age > 1 agent; every trigger consumes the list and runs
        k_epoch optimizer steps (PPO.py:136-162)."""
        w, lib, b = self.world, self.world.lib, self.brains[g]
        _lib.check(lib.rl_ppo_store(C.byref(w.cfg), C.byref(w.bufs), C.byref(self._rb_store), C.c_int32(g),
                                    C.c_void_p(self._prob.data_ptr()), C.c_int32(train_freq), C.byref(b._replay.bufs), st))
        self.gpu_launches += 3
        for _ in range(int(b.k_epoch)):
            _lib.check(lib.rl_ppo_epoch(C.byref(w.cfg), C.byref(self._rb_event), C.c_int32(g), C.byref(b._replay.bufs),
                                        C.byref(b._dev.learn_bufs), st))
            if self.dist:
                self._allreduce_grads([g])
            _lib.check(lib.rl_brain_adam(C.byref(b._dev.learn_bufs), st))
            self.gpu_launches += 6
        _lib.check(lib.rl_ppo_compact(C.byref(w.cfg), C.c_int32(g), C.byref(b._replay.bufs), st))
        self.gpu_launches += 1

    def _reduce_gate(self):
        if self.dist:
            self._gate.copy_(self.rows.total)
            torch.distributed.all_reduce(self._gate)

    def _allreduce_grads(self, active):
        """One NCCL all-reduce (sum) over the flattened gradient (+event count) of every active brain."""
        from ..sharding import allreduce_grads
        allreduce_grads([self.brains[g]._dev.grad for g in active])

    # ------------------------------------------------------------------ host views
    def agents_of(self, world=0):
        n = int(self.world.n_agents[world])
        rec = self.world.rec[world, :n].cpu().numpy().view(np.dtype(
            [("cell", "<u2"), ("health", "<i2"), ("age", "<i2"), ("max_age", "<i2"), ("gene", "<i4"), ("flags", "u1"),
             ("action", "i1"), ("prev_slot", "<u2")])).reshape(n)
        rew = self.world.reward[world, :n].cpu().numpy()
        return [AgentView(rec[s], self.width, float(rew[s]), self.brains[int(rec[s]["gene"])]) for s in range(n)]

    @property
    def agents(self):
        """World 0's agents in the reference's (row-major) order."""
        return self.agents_of(0)

    def check_status(self):
        """Raise what the reference would have raised inside the loop (device-side conditions are sticky flags, read at
        update_interval boundaries and at the end of trainer()): ValueError of random.sample on a buffer shorter than a
        batch (D3QN.py:140)."""
        self.sync_host_scalars()
        if int(self._sample_status) & 1:
            raise ValueError("Sample larger than population or is negative")
        for b in self.brains:
            st = getattr(getattr(b, "_replay", None), "status", None)
            if st is not None and int(st):
                raise RuntimeError(f"{b.method}: the per-world data list overflowed (status {int(st)}); raise data_capacity")
            mem = getattr(b, "memory", None)
            if mem is not None and int(mem.status):
                raise RuntimeError("PERDQN: a SumTree stratum found no filled leaf in 64 redraws (PERDQN.py:290-295 would "
                                   "keep drawing)")

    def sync_host_scalars(self):
        """Copy the device-side schedule state (epsilon, last n_epi seen) back into the brain objects (one small D2H)."""
        eps, seen = self._eps.cpu().tolist(), self._seen.cpu().tolist()
        for g, b in enumerate(self.brains):
            b._sync_host_scalars(eps[g], seen[g])

    def count_agents(self):
        """Total listed agents on this rank (device -> host read)."""
        return int(self.world.n_agents.sum())

    def epsilons(self):
        return self._eps.cpu().tolist()

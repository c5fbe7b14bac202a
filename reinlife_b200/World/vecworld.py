"""Device-resident batch of independent ReinLife worlds (SoA tensors in HBM) driven through the C ABI.

One `VecWorld` = `n_worlds` instances of the reference's `Environment` state
(ReinLife/World/environment.py:133-215), laid out as rl_world_bufs (include/reinlife_b200.h).
PyTorch is used for device memory and streams only.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib

REC_DTYPE = np.dtype([("cell", "<u2"), ("health", "<i2"), ("age", "<i2"), ("max_age", "<i2"),
                      ("gene", "<i4"), ("flags", "u1"), ("action", "i1"), ("prev_slot", "<u2")])


class VecWorld:
    def __init__(self, n_worlds, height=30, width=30, n_genes=2, max_agents=100, seed=0, world_id0=0,
                 slot_cap=None, static_families=True, limit_reproduction=False, incentivize_killing=True,
                 device="cuda", with_stats=False):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RLError("VecWorld needs a CUDA device: reinlife_b200 has no CPU path")
        self.n_worlds, self.H, self.W, self.G = n_worlds, height, width, n_genes
        self.C = height * width
        self.S = int(slot_cap or self.C)
        self.ld = _lib.OBS_LD
        self.cfg = _lib.WorldCfg(n_worlds, height, width, n_genes, max_agents, self.S, self.ld,
                                 int(static_families), int(limit_reproduction), int(incentivize_killing),
                                 seed, world_id0)
        dev = self.device
        self.type = torch.zeros((n_worlds, self.C), dtype=torch.uint8, device=dev)
        self.rec = torch.zeros((n_worlds, self.S, 16), dtype=torch.uint8, device=dev)
        self.n_agents = torch.zeros(n_worlds, dtype=torch.int32, device=dev)
        self.reward = torch.zeros((n_worlds, self.S), dtype=torch.float32, device=dev)
        self.obs_state = torch.zeros((n_worlds, self.S, self.ld), dtype=torch.float32, device=dev)
        self.obs_prime = torch.zeros((n_worlds, self.S, self.ld), dtype=torch.float32, device=dev)
        self.gene_count = torch.zeros((n_worlds, n_genes), dtype=torch.int32, device=dev)
        self.status = torch.zeros(n_worlds, dtype=torch.int32, device=dev)
        self.stats = torch.zeros((n_worlds, n_genes, _lib.N_STATS), dtype=torch.float32, device=dev) if with_stats else None
        self.reward_div100 = None      # allocated by enable_reward_div100() when a PPO brain trains (PPO.py:73)
        self.bufs = _lib.WorldBufs(self.type.data_ptr(), self.rec.data_ptr(), self.n_agents.data_ptr(),
                                   self.reward.data_ptr(), self.obs_state.data_ptr(), self.obs_prime.data_ptr(),
                                   self.gene_count.data_ptr(), self.status.data_ptr(),
                                   self.stats.data_ptr() if with_stats else None, None)
        self.t = 0
        # non-static families (rl_world_ns_bufs): Agent.fitness, object identity, max_gene / best agents / _produce event
        self.static_families = bool(static_families)
        self.ns = None
        if not self.static_families:
            self.fitness = torch.zeros((n_worlds, self.S), dtype=torch.float64, device=dev)
            self.serial = torch.zeros((n_worlds, self.S), dtype=torch.int64, device=dev)
            self.ns_state = torch.zeros((n_worlds, C.sizeof(_lib.NsState)), dtype=torch.uint8, device=dev)
            self.n_lineages = torch.zeros(n_worlds, dtype=torch.int32, device=dev)
            self.ns = _lib.WorldNsBufs(self.fitness.data_ptr(), self.serial.data_ptr(), self.ns_state.data_ptr(),
                                       self.n_lineages.data_ptr())

    def ns_host(self):
        """Host copy of the per-world non-static state as an array of _lib.NsState."""
        raw = self.ns_state.cpu().numpy().tobytes()
        return (_lib.NsState * self.n_worlds).from_buffer_copy(raw)

    def enable_obs_fp16(self):
        """Have the World kernels also write float16 copies of obs_state / obs_prime (column 159 = 1.0): the rows the tensor-core
        get_action kernel gathers by TMA and rl_replay_store copies into float16 rings."""
        if getattr(self, "obs_state_h", None) is None:
            if self.ld != 160:
                raise _lib.RLError("float16 observation copies need obs_ld = 160")
            self.obs_state_h = torch.zeros((self.n_worlds, self.S, 160), dtype=torch.float16, device=self.device)
            self.obs_prime_h = torch.zeros((self.n_worlds, self.S, 160), dtype=torch.float16, device=self.device)
            self.bufs.obs_state_h = self.obs_state_h.data_ptr()
            self.bufs.obs_prime_h = self.obs_prime_h.data_ptr()

    def enable_reward_div100(self):
        """Have rl_world_step also write float32(reward / 100.0), the value PPOAgent.learn stores (Models/PPO.py:73)."""
        if self.reward_div100 is None:
            self.reward_div100 = torch.zeros((self.n_worlds, self.S), dtype=torch.float32, device=self.device)
            self.bufs.reward_div100 = self.reward_div100.data_ptr()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # --- the reference's Environment phases -------------------------------------------------------
    def reset(self):
        with torch.cuda.device(self.device):
            if self.ns is not None:
                _lib.check(self.lib.rl_world_reset_ns(C.byref(self.cfg), C.byref(self.bufs), C.byref(self.ns), self._stream()))
            else:
                _lib.check(self.lib.rl_world_reset(C.byref(self.cfg), C.byref(self.bufs), self._stream()))
        self.t = 0

    def step(self):
        self.t += 1
        with torch.cuda.device(self.device):
            if self.ns is not None:
                _lib.check(self.lib.rl_world_step_ns(C.byref(self.cfg), C.byref(self.bufs), C.byref(self.ns), C.c_uint64(self.t), self._stream()))
            else:
                _lib.check(self.lib.rl_world_step(C.byref(self.cfg), C.byref(self.bufs), C.c_uint64(self.t), self._stream()))

    def update(self, top_up=None, max_age=50):
        """update_env; top_up=N additionally saturates the world to N agents in the same launch (static families only)."""
        with torch.cuda.device(self.device):
            if top_up and self.ns is None:
                _lib.check(self.lib.rl_world_update_top_up(C.byref(self.cfg), C.byref(self.bufs), C.c_uint64(self.t),
                                                           C.c_int32(top_up), C.c_int32(max_age), self._stream()))
                return
            if self.ns is not None:
                _lib.check(self.lib.rl_world_update_ns(C.byref(self.cfg), C.byref(self.bufs), C.byref(self.ns), C.c_uint64(self.t), self._stream()))
            else:
                _lib.check(self.lib.rl_world_update(C.byref(self.cfg), C.byref(self.bufs), C.c_uint64(self.t), self._stream()))

    def top_up(self, target, max_age=50):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rl_world_top_up(C.byref(self.cfg), C.byref(self.bufs), C.c_uint64(self.t),
                                                C.c_int32(target), C.c_int32(max_age), self._stream()))

    def observe(self, which=0):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rl_world_observe(C.byref(self.cfg), C.byref(self.bufs), C.c_int32(which), self._stream()))

    # --- host views (tests, trackers) -----------------------------------------------------------
    def rec_host(self):
        return self.rec.cpu().numpy().view(REC_DTYPE).reshape(self.n_worlds, self.S)

    def load_host(self, w, typ, rec):
        """Upload one world's canonical state (cell types + row-major agent list)."""
        self.type[w] = torch.from_numpy(np.ascontiguousarray(np.asarray(typ, np.uint8).reshape(-1))).to(self.device)
        n = len(rec)
        if n:
            raw = np.ascontiguousarray(rec).view(np.uint8).reshape(n, 16)
            self.rec[w, :n] = torch.from_numpy(raw.copy()).to(self.device)
        self.n_agents[w] = n

    def set_actions(self, actions):
        """actions: int8 tensor/array [n_worlds, <=S] in slot order -> rec[].action."""
        a = torch.as_tensor(actions, dtype=torch.int8, device=self.device)
        self.rec[:, :a.shape[1], 13] = a.view(torch.uint8)

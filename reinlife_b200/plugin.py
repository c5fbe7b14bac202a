"""The reference's per-agent `brain.learn(...)` plugin call with HOST buffers (World/entities.py:194-208 ->
Models/PERD3QN.py:117-125, D3QN.py:118-126, DQN.py:85-89, PPO.py:71-77, PERDQN.py:188-195).

A brain used on its own -- e.g. under the reference's own Environment / Agent classes -- owns a private one-world device
context: one replay ring (data list / SumTree memory) of the brain's capacity, its networks, Adam state.  Every
learn() call uploads the transition into slot 0 of that world and runs the SAME C-ABI sequence the vectorised
Environment runs, on one-row lists, in the reference's order of effects: store, then (on a trigger) sample -> train
-> priorities -> optimizer step(s), then the target-sync rule.  One call = a handful of tiny launches + a few hundred
bytes of H2D: slow by construction (the vectorised Environment.learn is the product path), exact by construction.
"""
import numpy as np
import torch

from . import _lib


# test hooks (tests/test_nonstatic_gpu.py): callables taking the brain -- ring positions used instead of the sampler's,
# and a callback after every train() event.  Module-level so that brains cloned mid-run are covered too.
SAMPLE_OVERRIDE = None
EVENT_HOOK = None


class PluginHost:
    def __init__(self, brain, device=None):
        from .World.environment import Environment
        if device is None:
            device = brain._dev.device if brain._dev is not None else torch.device("cuda", torch.cuda.current_device())
        self.brain = brain
        side = 3
        if brain.KIND == _lib.MODEL_PPO:       # the reference's data list is unbounded (PPO.py:113): room for 8192 transitions
            brain.data_capacity = int(brain.data_capacity or 8192)
            side = 46                          # PpoData plans min(capacity, 4 * slots) rows per step
        self.env = Environment(width=side, height=side, brains=[brain], max_agents=1, update_interval=10 ** 9, print_results=False,
                               training=True, n_worlds=1, seed=0, device=device, precision="fp32", sequential_events=True)
        w = self.env.world
        self.calls = 0
        self._rec = np.zeros(1, dtype=np.dtype([("cell", "<u2"), ("health", "<i2"), ("age", "<i2"), ("max_age", "<i2"),
                                                ("gene", "<i4"), ("flags", "u1"), ("action", "i1"), ("prev_slot", "<u2")]))
        w.n_agents[0] = 1
        # one-row STORE / EVENT lists of the only (world, brain): row id 0
        self.env.rows.rows.zero_()
        G3 = _lib.N_ROW_KINDS
        self._c1 = torch.ones(G3, dtype=torch.int32, device=w.device)
        self._c0 = torch.zeros(G3, dtype=torch.int32, device=w.device)
        self._gate = torch.zeros(G3, dtype=torch.int32, device=w.device)

    def learn(self, age, dead, action, state, reward, state_prime, done, n_epi=0, prob_a=None):
        env, b = self.env, self.brain
        w = env.world
        if bool(dead) != bool(done) and b.KIND == _lib.MODEL_PPO:
            raise ValueError("PPO plugin learn(): `dead` and `done` must agree (they are the same flag in World/entities.py:194-208)")
        state = np.asarray(state, np.float64).reshape(-1)
        state_prime = np.asarray(state_prime, np.float64).reshape(-1)
        if state.shape[0] != _lib.OBS_DIM or state_prime.shape[0] != _lib.OBS_DIM:
            raise ValueError(f"expected {_lib.OBS_DIM}-value observations")
        r = self._rec
        r["age"], r["max_age"], r["health"], r["gene"] = int(age), 32767, 1, 0
        r["flags"] = _lib.F_DEAD if done else 0
        r["action"], r["prev_slot"] = int(action), 0
        rows = np.zeros((2, w.ld), np.float32)
        rows[0, :_lib.OBS_DIM], rows[1, :_lib.OBS_DIM] = state, state_prime           # torch.FloatTensor(obs) rounding
        w.rec[0, 0].copy_(torch.from_numpy(r.view(np.uint8).reshape(16).copy()))
        both = torch.from_numpy(rows).to(w.device)
        w.obs_state[0, 0].copy_(both[0]); w.obs_prime[0, 0].copy_(both[1])
        w.reward[0, 0] = float(np.float32(reward))
        if w.reward_div100 is not None:
            w.reward_div100[0, 0] = float(np.float32(float(reward) / 100.0))      # PPO.py:73: reward / 100.0 in float64
        if prob_a is not None:
            env._prob[0] = float(prob_a)
        tf = [int(getattr(b, "train_freq", 1))]
        on = [int(n_epi > getattr(b, "exploration", -1))]
        trigger = bool(on[0]) and (int(age) % tf[0] == 0 or bool(dead))
        self.calls += 1
        # the one-row views of Environment._one_row_views, on constant lists (row id 0 for every kind)
        rb = env.rows.bufs
        c1, c0 = self._c1.data_ptr(), self._c0.data_ptr()
        env._rb_store = _lib.RowsBufs(c1, c0, c1, rb.rows, rb.row_cap, 0)
        ce = c1 if trigger else c0
        env._rb_event = _lib.RowsBufs(ce, c0, ce, rb.rows, rb.row_cap, 0)
        self._gate[_lib.ROWS_STORE] = 1
        self._gate[_lib.ROWS_EVENT] = 1 if trigger else 0
        env._gate_base = self._gate.data_ptr()
        env._t_key = self.calls
        if SAMPLE_OVERRIDE is not None:
            env.sample_override = lambda _g, _k: SAMPLE_OVERRIDE(b)
        if EVENT_HOOK is not None:
            env.event_hook = lambda _g, _k: EVENT_HOOK(b)
        env._seq_event = (0, 0) if trigger else None
        env._learn_lists([0], tf, on, n_epi)
        env._seq_event = None
        if trigger and env.event_hook is not None:
            env.event_hook(0, 0)
        self._raise_status()

    def _raise_status(self):
        """Device-side error flags, raised where the reference raises (inside the call).  The epsilon of the host-driven
        schedules stays host-owned here (get_action decays brain.epsilon like the reference); only PERDQN's, which moves
        inside train_model (PERDQN.py:132-133), is read back from the device."""
        env, b = self.env, self.brain
        if int(env._sample_status) & 1:
            env._sample_status.zero_()
            raise ValueError("Sample larger than population or is negative")          # random.sample, D3QN.py:140
        st = getattr(getattr(b, "_replay", None), "status", None)
        if st is not None and int(st):
            raise RuntimeError(f"{b.method}: the data list overflowed (status {int(st)}); raise data_capacity")
        mem = getattr(b, "memory", None)
        if mem is not None:
            if int(mem.status):
                raise RuntimeError("PERDQN: a SumTree stratum found no filled leaf in 64 redraws")
            b.epsilon = float(env._eps[0])

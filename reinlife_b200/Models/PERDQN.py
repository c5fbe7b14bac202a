"""PERDQN brain -- same constructor, attributes and method string as ReinLife/Models/PERDQN.py:14-91: 153-64-64-8 MLP
(xavier weights), SumTree prioritized memory, epsilon 1.0 -> 0.01 stepping once per train_model, importance-weighted
MSE, target copied at every learn trigger.  Network, optimizer and the per-world memories live on the device once
bound to an Environment (kernels: csrc/sumtree_kernels.cu, csrc/learn_rows_kernels.cu)."""
import random

import numpy as np
import torch

from .. import _lib
from . import packing
from ._base import DeviceBrainBase, _NetHandle


class PERDQNAgent(DeviceBrainBase):
    KIND, RULE, PRIORITIZED, HAS_TARGET = packing.PERDQN, _lib.ACT_PERDQN, False, True
    DEVICE_LEARN = True

    def __init__(self, input_dim=153, output_dim=8, explore_step=5_000, train_freq=20, learning_rate=0.001,
                 batch_size=64, gamma=0.99, capacity=20000, load_model=False, training=True):
        super().__init__(input_dim, output_dim, "PERDQN")
        if input_dim != 153 or output_dim != 8:
            raise ValueError("the device brains are specialised for ReinLife's 153-float observation and 8 actions")
        if batch_size != 64:
            raise ValueError("batch_size must be 64 (one train_model() event = one 64-row tile)")
        self._init_common()
        self.state_size, self.action_size = input_dim, output_dim
        self.discount_factor = gamma
        self.learning_rate = learning_rate
        self.memory_size = capacity            # one Memory of this capacity per (world, brain)
        self.epsilon = 1.0
        self.epsilon_min = 0.01
        self.explore_step = explore_step
        self.epsilon_decay = (self.epsilon - self.epsilon_min) / self.explore_step
        self.batch_size = batch_size
        self.train_start = 1000
        self.training = training
        self.train_freq = train_freq
        self.memory = None                     # brains.SumTrees once bound to a training Environment
        # reference construction order (PERDQN.py:73-83): model (+ xavier), target_model (default draw, then overwritten
        # by update_target_model)
        self._host_sd = packing.default_init(self.KIND)
        self._burn_target_draw()
        self._host_sd_target = {k: v.clone() for k, v in self._host_sd.items()}
        self.model = _NetHandle(self)
        self.target_model = _NetHandle(self, target=True)
        if not self.training:
            self.epsilon = 0
        if load_model:                         # PERDQN.py:88-90: the online model only
            self.model.load_state_dict(torch.load(load_model, map_location="cpu"))

    @staticmethod
    def _burn_target_draw():
        """target_model = DQN(...) consumes torch's global RNG like the reference (default nn.Linear init, no xavier)."""
        import torch.nn as nn
        for i, o in ((153, 64), (64, 64), (64, 8)):
            nn.Linear(i, o)

    def _lr(self): return self.learning_rate
    def _gamma(self): return self.discount_factor
    def _batch(self): return self.batch_size
    def _capacity(self): return self.memory_size

    def _bind(self, env, gene):
        from ..brains import ReplayRings, SumTrees
        if env.training and self._trains() and self.memory is None:
            need = ReplayRings.bytes_needed(env.n_worlds, self.memory_size) + SumTrees.bytes_needed(env.n_worlds, self.memory_size)
            free, _ = torch.cuda.mem_get_info(env.device)
            if need > 0.9 * free:
                raise MemoryError(f"PERDQN memories for {env.n_worlds} worlds x capacity {self.memory_size} need "
                                  f"{need / 2**30:.1f} GiB, {free / 2**30:.1f} GiB free; lower `capacity`")
        super()._bind(env, gene)
        if env.training and self._trains() and self.memory is None:
            self.memory = SumTrees(env.n_worlds, self.memory_size, env.device, train_start=self.train_start,
                                   ev_cap=env.rows.row_cap)

    def update_target_model(self):             # PERDQN.py:97-99
        self.target_model.load_state_dict(self.model.state_dict())

    # ---- the reference's per-agent plugin calls (host observations) ---------------------------------
    def get_action(self, state):               # PERDQN.py:101-111
        if np.random.rand() <= self.epsilon:
            return random.randrange(self.action_size)
        q = self._q_single(np.asarray(state))
        return int(np.argmax(q))

    def learn(self, age, dead, action, state, reward, state_prime, done):
        """PERDQN.py:188-195: append_sample; on a trigger train_model() once the memory holds train_start items, then
        target_model <- model."""
        self._plugin_learn(age=age, dead=dead, action=action, state=state, reward=reward, state_prime=state_prime, done=done)

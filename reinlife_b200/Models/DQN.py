"""DQN brain (ReinLife/Models/DQN.py:19-89): 153-128-64-8 MLP, linear epsilon, uniform replay of 50 000."""
import random

import numpy as np
import torch

from .. import _lib
from . import packing
from ._base import DeviceBrainBase, _NetHandle

gamma = 0.98
buffer_limit = 50000
batch_size = 32


class DQNAgent(DeviceBrainBase):
    KIND, RULE, PRIORITIZED, HAS_TARGET = packing.DQN, _lib.ACT_DQN, False, True
    DEVICE_LEARN = True

    def __init__(self, input_dim=153, output_dim=8, max_epi=0, learning_rate=0.0005, train_freq=20,
                 load_model=False, training=True, *, buffer_limit=buffer_limit):
        super().__init__(input_dim, output_dim, "DQN")
        if input_dim != 153 or output_dim != 8:
            raise ValueError("the device brains are specialised for ReinLife's 153-float observation and 8 actions")
        self._init_common()
        self._host_sd = packing.default_init(self.KIND)     # agent (DQN.py:48)
        packing.default_init(self.KIND)                     # target draw, overwritten by load_state_dict (:49-50)
        self._host_sd_target = {k: v.clone() for k, v in self._host_sd.items()}
        self.agent = _NetHandle(self)
        self.target = _NetHandle(self, target=True)
        self.learning_rate = learning_rate
        self.max_epi = max_epi
        self.epsilon = 0.20
        self.train_freq = train_freq
        self.training = training
        # module constant in the reference (DQN.py:15: deque(maxlen=50000) per brain); one ring per (world, brain) here,
        # so many-world runs pass a smaller keyword-only `buffer_limit`.  train() needs size() > 1000 (DQN.py:79).
        self.buffer_limit = int(buffer_limit)
        self.min_buffer = 1000
        if not self.training:
            self.epsilon = 0
        if load_model:
            self.agent.load_state_dict(torch.load(load_model, map_location="cpu"))

    def _lr(self): return self.learning_rate
    def _gamma(self): return gamma
    def _batch(self): return batch_size
    def _capacity(self): return self.buffer_limit

    def _sched(self):
        if self.training and self.max_epi == 0:
            # the reference divides by max_epi at n_epi % 30 == 0 (DQN.py:69): ZeroDivisionError on the first step
            raise ZeroDivisionError("float division by zero (DQN(max_epi=0) while training, DQN.py:69)")
        return _lib.BrainSched(self.RULE, int(bool(self.training)), 0.01, 1.0, int(self.max_epi))

    def get_action(self, state, n_epi):            # DQN.py:65-71, 132-139
        if self.training:
            if n_epi % 30 == 0:
                self.epsilon = max(0.01, 0.20 - 0.20 * (n_epi / self.max_epi))
        q = self._q_single(np.asarray(state))
        coin = random.random()
        if coin < self.epsilon:
            return random.randint(0, 7)
        return int(np.argmax(q))

    def learn(self, age, dead, action, state, reward, state_prime, done):
        """DQN.py:85-89: memorize; on a trigger train() (5 x sample 32 once the buffer holds > 1000) and target <- agent."""
        self._plugin_learn(age=age, dead=dead, action=action, state=state, reward=reward, state_prime=state_prime, done=done)

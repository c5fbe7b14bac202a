"""PERD3QN brain -- same constructor, attributes and method string as ReinLife/Models/PERD3QN.py:10-79;
the networks, optimizer and prioritized replay live on the device once bound to an Environment."""
import random

import numpy as np
import torch

from .. import _lib
from . import packing
from ._base import DeviceBrainBase, _NetHandle


class PERD3QNAgent(DeviceBrainBase):
    KIND, RULE, PRIORITIZED, HAS_TARGET = packing.DUELING, _lib.ACT_DUELING, True, True
    METHOD = "PERD3QN"
    DEVICE_LEARN = True

    def __init__(self, input_dim=153, output_dim=8, exploration=1000, soft_update_freq=200, train_freq=20,
                 learning_rate=1e-3, batch_size=64, capacity=10000, gamma=0.99, load_model=False, training=True):
        super().__init__(input_dim, output_dim, self.METHOD)
        if input_dim != 153 or output_dim != 8:
            raise ValueError("the device brains are specialised for ReinLife's 153-float observation and 8 actions")
        if batch_size != 64:
            raise ValueError("batch_size must be 64 (one train() event = one 64-row tile)")
        self._init_common()
        # reference construction order (PERD3QN.py:51-53): target_net, eval_net, eval <- target
        self._host_sd = packing.default_init(self.KIND)
        packing.default_init(self.KIND)           # the eval_net draw the reference throws away
        self._host_sd_target = {k: v.clone() for k, v in self._host_sd.items()}
        self.target_net = _NetHandle(self, target=True)
        self.eval_net = _NetHandle(self)
        self.learning_rate = learning_rate
        self.capacity = capacity
        self.exploration = exploration
        self.soft_update_freq = soft_update_freq
        self.train_freq = train_freq
        self.batch_size = batch_size
        self.gamma = gamma
        self.n_epi = 0
        self.epsilon = 0.9
        self.epsilon_min = 0.05
        self.decay = 0.99
        self.training = training
        if not self.training:
            self.epsilon = 0
        if load_model:                             # PERD3QN.py:72-79
            sd = torch.load(load_model, map_location="cpu")
            self.eval_net.load_state_dict(sd)
            if self.training:
                self.target_net.load_state_dict(sd)

    def _lr(self): return self.learning_rate
    def _gamma(self): return self.gamma
    def _batch(self): return self.batch_size
    def _capacity(self): return self.capacity

    def _sched(self):
        return _lib.BrainSched(self.RULE, int(bool(self.training)), self.epsilon_min, self.decay, 0)

    # ---- the reference's per-agent plugin calls (host observations) ---------------------------------
    def get_action(self, state, n_epi):            # PERD3QN.py:81-89, 204-210
        if self.training:
            if n_epi > self.n_epi:
                if self.epsilon > self.epsilon_min:
                    self.epsilon = self.epsilon * self.decay
                self.n_epi = n_epi
        if random.random() > self.epsilon:
            q = self._q_single(np.asarray(state))
            return int(np.argmax(q))
        return random.choice(list(range(self.output_dim)))

    def learn(self, age, dead, action, state, reward, state_prime, done, n_epi):
        """PERD3QN.py:117-125 / D3QN.py:118-126: memorize; if n_epi > exploration: train() on a trigger, target sync every
        soft_update_freq episodes -- on this brain's private device ring (reinlife_b200/plugin.py)."""
        self._plugin_learn(age=age, dead=dead, action=action, state=state, reward=reward, state_prime=state_prime,
                           done=done, n_epi=n_epi)

    def apply_gaussian_noise(self):                # PERD3QN.py:127-130: net effect = target <- eval
        self.target_net.load_state_dict(self.eval_net.state_dict())

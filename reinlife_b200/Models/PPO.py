"""PPO brain (ReinLife/Models/PPO.py:10-77): 153-256-256 trunk, softmax policy head + value head."""
import numpy as np
import torch

from .. import _lib
from . import packing
from ._base import DeviceBrainBase, _NetHandle


class PPOAgent(DeviceBrainBase):
    KIND, RULE, PRIORITIZED, HAS_TARGET = packing.PPO, _lib.ACT_PPO, False, False
    DEVICE_LEARN = True

    def __init__(self, input_dim=153, output_dim=8, learning_rate=0.0005, gamma=0.98, lmbda=0.95, eps_clip=0.1,
                 k_epoch=3, train_freq=20, load_model=False, *, data_capacity=None):
        super().__init__(input_dim, output_dim, "PPO")
        if input_dim != 153 or output_dim != 8:
            raise ValueError("the device brains are specialised for ReinLife's 153-float observation and 8 actions")
        self._init_common()
        self._host_sd = packing.default_init(self.KIND)
        self._host_sd_target = None
        self.model = _NetHandle(self)
        self.learning_rate, self.gamma, self.lmbda, self.eps_clip, self.k_epoch = learning_rate, gamma, lmbda, eps_clip, k_epoch
        self.load_model = load_model
        self.train_freq = train_freq
        # the reference's `data` list is unbounded (PPO.py:113-115); here one fixed-size list per (world, brain):
        # rows per world, default 32 x max_agents (a list is consumed at every train trigger, PPO.py:75-77,133)
        self.data_capacity = data_capacity
        self.training = False if load_model else True
        if self.load_model:
            self.model.load_state_dict(torch.load(load_model, map_location="cpu"))

    def _lr(self): return self.learning_rate
    def _gamma(self): return self.gamma
    def _batch(self): return 64
    def _capacity(self): return self.data_capacity

    def _bind(self, env, gene):
        from ..brains import DeviceBrain, PpoData
        if self._env is not None and self._env is not env:
            raise RuntimeError("a brain object can be bound to one Environment only")
        if self._dev is None:
            self._dev = DeviceBrain(self.KIND, self._host_sd, env.device, lr=self.learning_rate, gamma=self.gamma,
                                    batch=64, has_target=False)
        self._env, self._gene = env, gene
        if env.training and self._trains():
            cap = int(self.data_capacity or min(8192, 32 * env.max_agents))
            if self._replay is None:
                row_cap = env.n_worlds * min(cap, 4 * env.world.S)
                self._replay = PpoData(env.n_worlds, cap, row_cap, env.device, lmbda=self.lmbda, eps_clip=self.eps_clip)
            env.world.enable_reward_div100()
            self._dev.alloc_learn(env.rows.row_cap, need_batch_bufs=False)

    def get_action(self, s):                       # PPO.py:54-60, 164-169
        prob = torch.from_numpy(np.asarray(self._q_single(np.asarray(s)), np.float32))
        a = int(torch.distributions.Categorical(prob).sample().item())
        return a if self.load_model else (a, prob)

    def learn(self, age, dead, action, state, reward, state_prime, done, prob):
        """PPO.py:71-77: put_data((state, action, reward / 100.0, state_prime, prob[action].item(), done)); on a trigger
        learn() consumes the data list (k_epoch optimizer steps)."""
        self._plugin_learn(age=age, dead=dead, action=action, state=state, reward=reward, state_prime=state_prime, done=done,
                           prob_a=float(prob[int(action)]))

"""D3QN brain (ReinLife/Models/D3QN.py:16-80): the PERD3QN network and schedule with a uniform replay buffer."""
from .PERD3QN import PERD3QNAgent


class D3QNAgent(PERD3QNAgent):
    PRIORITIZED = False
    METHOD = "D3QN"
    DEVICE_LEARN = True    # uniform random.sample replay (D3QN.py:138-142) = rl_replay_sample_uniform

    def __init__(self, input_dim=153, output_dim=8, exploration=1000, soft_update_freq=200, train_freq=20,
                 learning_rate=1e-3, gamma=0.99, batch_size=64, capacity=10000, load_model=False, training=True):
        super().__init__(input_dim, output_dim, exploration, soft_update_freq, train_freq, learning_rate, batch_size,
                         capacity, gamma, load_model, training)

    def apply_gaussian_noise(self):
        raise AttributeError("D3QN has no apply_gaussian_noise (reference: only PERD3QN, World/entities.py:210-213)")

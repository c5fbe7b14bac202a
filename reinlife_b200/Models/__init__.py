"""Same export list as ReinLife/Models/__init__.py:1-5."""
from .D3QN import D3QNAgent as D3QN          # noqa: F401
from .DQN import DQNAgent as DQN             # noqa: F401
from .PERDQN import PERDQNAgent as PERDQN    # noqa: F401
from .PERD3QN import PERD3QNAgent as PERD3QN  # noqa: F401
from .PPO import PPOAgent as PPO             # noqa: F401

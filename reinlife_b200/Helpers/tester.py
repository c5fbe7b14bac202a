"""tester() -- ReinLife/Helpers/tester.py:6-72 without the pygame window.  The reference loops forever; here
`n_steps` bounds the loop (None = forever, like the reference) and the environment is returned."""
from typing import List

from ..World.environment import Environment


def tester(brains: List, width: int = 30, height: int = 30, max_agents: int = 100, pastel_colors: bool = False,
           static_families: bool = True, limit_reproduction: bool = False, fps: int = 10, *, n_worlds: int = 1,
           seed: int = 0, device=None, n_steps=None, saturate_to: int = 0) -> Environment:
    env = Environment(width=width, height=height, grid_size=24, max_agents=max_agents, pastel_colors=pastel_colors,
                      brains=brains, training=False, static_families=static_families,
                      limit_reproduction=limit_reproduction, n_worlds=n_worlds, seed=seed, device=device)
    env.reset()
    if saturate_to:
        env.top_up(saturate_to)
    k = 0
    while n_steps is None or k < n_steps:
        env.act(0)                # agent.action = agent.brain.get_action(agent.state[, 0])   tester.py:58-68
        env.step()
        if saturate_to:
            env.update_env(top_up=saturate_to)
        else:
            env.update_env()
        k += 1
    return env

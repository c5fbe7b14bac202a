"""trainer() -- the reference's public training entry point (ReinLife/Helpers/trainer.py:7-107), same signature and
loop order (get_action -> env.step -> learn -> env.update_env), the two per-agent loops replaced by the batched
Environment.act / Environment.learn.  New optional keywords: n_worlds, seed, device, saturate_to, precision,
sequential_events (n_worlds=1: the reference's exact per-agent order of learning effects).  static_families=False runs the
evolving-lineage mode (reinlife_b200/World/nonstatic.py)."""
from typing import List

from ..World.environment import Environment


def trainer(brains: List, n_episodes: int = 10_000, width: int = 30, height: int = 30,
            visualize_results: bool = False, google_colab: bool = False, update_interval: int = 500,
            print_results: bool = True, max_agents: int = 100, render: bool = False, static_families: bool = True,
            training: bool = True, save: bool = True, limit_reproduction: bool = False,
            incentivize_killing: bool = True, *, n_worlds: int = 1, seed: int = 0, device=None,
            saturate_to: int = 0, precision: str = "fp16", sequential_events: bool = False) -> Environment:
    env = Environment(width=width, height=height, max_agents=max_agents, brains=brains, grid_size=24,
                      static_families=static_families, update_interval=update_interval, print_results=print_results,
                      interactive_results=visualize_results, google_colab=google_colab, training=training,
                      limit_reproduction=limit_reproduction, incentivize_killing=incentivize_killing,
                      n_worlds=n_worlds, seed=seed, device=device, precision=precision,
                      sequential_events=sequential_events)
    env.reset()
    if saturate_to and not static_families:
        raise ValueError("saturate_to (the benchmark's saturated-world generator) is defined for static families only")
    if saturate_to:
        env.top_up(saturate_to)
    if render:
        raise NotImplementedError("render=True needs pygame; the renderer is out of scope")

    for n_epi in range(n_episodes + 1):
        env.act(n_epi)            # for agent in env.agents: agent.get_action(n_epi)      trainer.py:88-89
        env.step()                #                                                        trainer.py:92
        env.learn(n_epi)          # for agent in env.agents: agent.learn(n_epi=n_epi)     trainer.py:95-96
        if saturate_to:           # trainer.py:99 + the benchmark's saturated-world generator in the same launch
            env.update_env(n_epi, top_up=saturate_to)
        else:
            env.update_env(n_epi)

    env.check_status()
    if save:
        env.save_results()
    return env

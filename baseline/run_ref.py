"""Times the UNMODIFIED Python reference (MaartenGr/ReinLife, installed by
`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`, git-ignored,
shipped to the GPU box with the gpurun snapshot) on host cores -- BASELINE.md section 3, scenario SAT-train:

    one 30x30 world, saturated to 100 agents (topped up after every update_env with the harness below, top-up time
    excluded), brains = [PERD3QNAgent(exploration=0), PERD3QNAgent(exploration=0)], the reference's own loop body
    (Helpers/trainer.py:85-99): get_action for every agent -> env.step() -> learn for every agent -> env.update_env().

Nothing of reinlife_b200 / oracle is imported here: this is the reference's stock code path through its public classes.
pygame / matplotlib are absent from the image and only used by the renderer / plots: stubbed in sys.modules.

    python baseline/run_ref.py --procs 1 --steps 60 --warmup 10            # one process
    python baseline/run_ref.py --procs 0 --steps 60 --warmup 10            # one process per usable core (aggregate)
"""
import argparse
import json
import os
import subprocess
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "ReinLife"))


def _load():
    for name in ("pygame", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.pyplot"].Figure = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import ReinLife  # noqa: F401
    return ReinLife


def _top_up(env, target, rng):
    """Saturated-world generator of SURVEY.md 8d with the reference's own objects: agents on uniformly random empty cells
    until `target` are alive, gene uniform over the brains, health 10*U{1..20}, age U{0..max_age-1}; then re-observe."""
    import numpy as np
    from ReinLife.World.entities import Agent
    n = len(env.grid.get_entities(env.entities.agent))
    added = False
    while n < target:
        ii, jj = np.where(env.grid.get_numpy() == 0)
        if len(ii) == 0:
            break
        e = int(rng.integers(len(ii)))
        g = int(rng.integers(len(env.brains)))
        a = env.grid.set(int(ii[e]), int(jj[e]), Agent, brain=env.brains[g], gene=g)
        a.health = 10 * int(rng.integers(1, 21))
        a.age = int(rng.integers(0, 50))
        n += 1
        added = True
    if added:
        env._get_observations()
        env._update_agents_state()


def worker(seed, steps, warmup, training=True):
    import warnings
    warnings.filterwarnings("ignore")
    import random
    import numpy as np
    import torch
    torch.set_num_threads(1)
    pkg = _load()
    from ReinLife.Models import PERD3QN
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    brains = [PERD3QN(exploration=0), PERD3QN(exploration=0)]
    env = pkg.Environment(width=30, height=30, brains=brains, grid_size=24, max_agents=100, update_interval=10 ** 9,
                          print_results=False, static_families=True, training=training)
    env.reset()
    _top_up(env, 100, rng)
    agent_steps, sec = 0, 0.0
    for n_epi in range(1, warmup + steps + 1):
        t0 = time.perf_counter()
        n = len(env.agents)
        for agent in env.agents:                     # Helpers/trainer.py:88-89
            agent.get_action(n_epi)
        env.step()                                   # :92
        if training:
            for agent in env.agents:                 # :95-96
                agent.learn(n_epi=n_epi)
        env.update_env(n_epi)                        # :99
        dt = time.perf_counter() - t0
        _top_up(env, 100, rng)                       # excluded from the timing
        if n_epi > warmup:
            agent_steps += n
            sec += dt
    return agent_steps, sec


def run_parallel(procs, steps, warmup, timeout=1500):
    """-> (aggregate agent*steps/s, processes used, per-process list).  procs <= 0: one per usable core."""
    if procs <= 0:
        try:
            procs = len(os.sched_getaffinity(0))
        except AttributeError:
            procs = os.cpu_count() or 1
        procs = max(1, min(procs, 128))
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", "--seed", str(s + 1), "--steps", str(steps),
                            "--warmup", str(warmup)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
          for s in range(procs)]
    res = []
    for p in ps:
        out, err = p.communicate(timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError(f"reference worker failed: {err[-800:]}")
        res.append(json.loads(out.strip().splitlines()[-1]))
    # every process runs the same number of steps concurrently: aggregate = sum of the per-process rates
    value = sum(r["agent_steps"] / r["sec"] for r in res)
    return value, procs, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--worker", action="store_true")
    a = ap.parse_args()
    if not available():
        print(json.dumps({"unavailable": "baseline/_ref is missing (pip install --target baseline/_ref of the reference)"}))
        return
    if a.worker:
        n, sec = worker(a.seed, a.steps, a.warmup)
        print(json.dumps({"agent_steps": n, "sec": sec, "seed": a.seed}))
        return
    value, procs, res = run_parallel(a.procs, a.steps, a.warmup)
    print(json.dumps({"agent_steps_per_sec": value, "procs": procs, "steps": a.steps, "warmup": a.warmup,
                      "per_process": [r["agent_steps"] / r["sec"] for r in res]}))


if __name__ == "__main__":
    main()

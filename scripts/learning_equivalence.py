"""Learning-outcome evidence for the reduced-precision event kernels: the same 256-world x 2000-step
PERD3QN x2 training run (natural worlds, exploration=200, train_freq=20, capacity=2000, update_interval=100) under
precision = fp32 (CUDA-core FMA, the reference arithmetic), tf32 and fp16 (tensor cores), six seeds each (initial weights and
world seeds vary together).  The two brains COMPETE, so which gene ends up dominant is a symmetry breaking that any
perturbation flips; the comparison is therefore on gene-symmetric summaries (total population, dominant share,
population-weighted / dominant / subordinate fitness, best age, attack ratio), mean +- std over the seeds.

    python scripts/learning_equivalence.py > profiles/learning_equivalence_r02.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reinlife_b200 as rl                      # noqa: E402
from reinlife_b200.Models import PERD3QN        # noqa: E402

N_WORLDS, N_EPI, INTERVAL = 256, 2000, 100
KEYS = ["Avg Population Size", "Avg Population Age", "Avg Population Fitness", "Best Population Age", "Avg Number of Attacks"]


def run(precision, seed):
    torch.manual_seed(1000 + seed)               # initial weights AND world seed vary with `seed`
    brains = [PERD3QN(exploration=200, capacity=2000), PERD3QN(exploration=200, capacity=2000)]
    t0 = time.time()
    env = rl.trainer(brains, n_episodes=N_EPI, width=30, height=30, max_agents=100, update_interval=INTERVAL, print_results=False,
                     save=False, n_worlds=N_WORLDS, seed=seed, precision=precision)
    torch.cuda.synchronize()
    res = env.tracker.results
    late = {k: [float(np.nanmean(res[k][g][-5:])) for g in sorted(res[k])] for k in KEYS}
    pop = late["Avg Population Size"]
    dom = int(np.argmax(pop))                    # the two brains compete: which gene ends up dominant is a symmetry breaking
    w = np.array(pop) / sum(pop)
    summary = {"total_population": float(sum(pop)), "dominant_share": float(max(w)),
               "fitness_population_weighted": float((w * np.array(late["Avg Population Fitness"])).sum()),
               "fitness_dominant": late["Avg Population Fitness"][dom], "fitness_subordinate": late["Avg Population Fitness"][1 - dom],
               "best_age_mean": float(np.mean(late["Best Population Age"])), "attack_ratio_mean": float(np.mean(late["Avg Number of Attacks"]))}
    return {"precision": precision, "seed": seed, "sec": time.time() - t0, "adam_steps": [int(b._dev.adam_step) for b in brains],
            "late_mean": late, "summary": summary,
            "curves": {k: {str(g): res[k][g] for g in res[k]} for k in ("Avg Population Size", "Avg Population Fitness")}}


if __name__ == "__main__":
    seeds = [0, 1, 2, 3, 4, 5]
    runs = [run(p, sd) for p in ("fp32", "tf32", "fp16") for sd in seeds]
    stats = {}
    for p in ("fp32", "tf32", "fp16"):
        rs = [r["summary"] for r in runs if r["precision"] == p]
        stats[p] = {k: {"mean": float(np.mean([x[k] for x in rs])), "std": float(np.std([x[k] for x in rs], ddof=1))} for k in rs[0]}
    print(json.dumps({"config": {"n_worlds": N_WORLDS, "n_episodes": N_EPI, "update_interval": INTERVAL, "seeds": seeds,
                                 "brains": "PERD3QN x2, exploration=200, capacity=2000, train_freq=20; late mean = last 5 intervals"},
                      "summary_mean_std_over_seeds": stats, "runs": runs}, indent=1))

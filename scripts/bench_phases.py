"""Phase times of the benchmark loop (act/step/learn/update/top_up) in context, few steps; for A/B runs of library builds."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import reinlife_b200 as rl
from reinlife_b200.Models import PERD3QN
torch.manual_seed(0)
brains = [PERD3QN(exploration=0, capacity=2000), PERD3QN(exploration=0, capacity=2000)]
env = rl.Environment(width=30, height=30, brains=brains, max_agents=100, print_results=False, training=True, n_worlds=4096, seed=0, precision=os.environ.get("RL_PRECISION", "fp16"))
env.reset(); env.top_up(100)
names = ["act", "step", "learn", "update", "top_up"]
tot = {k: 0.0 for k in names}
n = 0
for n_epi in range(1, 19):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record(); env.act(n_epi); ev[1].record(); env.step(); ev[2].record(); env.learn(n_epi); ev[3].record()
    env.update_env(n_epi); ev[4].record(); env.top_up(100); ev[5].record()
    torch.cuda.synchronize()
    if n_epi > 6:
        n += 1
        for k, name in enumerate(names):
            tot[name] += ev[k].elapsed_time(ev[k + 1])
print(os.environ.get("REINLIFE_B200_LIB", "default").split("/")[-1], {k: round(v / n, 4) for k, v in tot.items()})

"""Throughput of the other BASELINE.json configurations (not bench.py lines; numbers quoted in DESIGN.md section 6).

    python scripts/bench_configs.py tester_dqn      # configs[1]: 256 worlds, 30x30, DQN x1, inference only (tester loop body)
    python scripts/bench_configs.py ppo_perd3qn     # configs[4] brain mix on 60x60 / 400 agents, 1024 worlds, STATIC families
    python scripts/bench_configs.py d3qn | dqn | ppo | perdqn  # configs[2] shape with the other trainable families (fp32 learn kernels)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import reinlife_b200 as rl
from reinlife_b200.Models import D3QN, DQN, PERD3QN, PERDQN, PPO

which = sys.argv[1] if len(sys.argv) > 1 else "tester_dqn"
steps, warm = 30, 8
torch.manual_seed(0)
if which == "tester_dqn":
    brains, kw, training = [DQN(training=False)], dict(width=30, height=30, max_agents=100, n_worlds=256), False
elif which == "ppo_perd3qn":
    brains = [PPO(), PERD3QN(exploration=0, capacity=2000)]
    kw, training = dict(width=60, height=60, max_agents=400, n_worlds=1024), True
elif which == "d3qn":
    brains = [D3QN(exploration=3, capacity=2000), D3QN(exploration=3, capacity=2000)]   # >= 64 items before the first train event (D3QN.py:140)
    kw, training = dict(width=30, height=30, max_agents=100, n_worlds=4096), True
    warm = 12                                  # the deque must hold >= 64 items before the first train event
elif which == "dqn":
    brains = [DQN(max_epi=1000, buffer_limit=2000), DQN(max_epi=1000, buffer_limit=2000)]
    kw, training = dict(width=30, height=30, max_agents=100, n_worlds=4096), True
    warm = 40                                  # rings must hold > 1000 items before train() does anything (DQN.py:79)
elif which == "perdqn":
    brains = [PERDQN(capacity=2000), PERDQN(capacity=2000)]
    kw, training = dict(width=30, height=30, max_agents=100, n_worlds=4096), True
    warm = 30                                  # memories must hold >= train_start = 1000 items (PERDQN.py:192): ~45 stores/step
elif which == "ppo":
    brains, kw, training = [PPO(), PPO()], dict(width=30, height=30, max_agents=100, n_worlds=4096), True
else:
    raise SystemExit(__doc__)
target = kw["max_agents"]
env = rl.Environment(brains=brains, print_results=False, training=training, seed=0, update_interval=10_000, **kw)
env.reset(); env.top_up(target)
count = torch.zeros(1, dtype=torch.int64, device=env.device)
names = ["act", "step", "learn", "update", "top_up"]
tot = {k: 0.0 for k in names}


def body(n_epi, timed):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    count.add_(env.world.n_agents.sum())
    ev[0].record(); env.act(n_epi); ev[1].record(); env.step(); ev[2].record()
    if training:
        env.learn(n_epi)
    ev[3].record(); env.update_env(n_epi); ev[4].record(); env.top_up(target); ev[5].record()
    if timed:
        torch.cuda.synchronize()
        for k, name in enumerate(names):
            tot[name] += ev[k].elapsed_time(ev[k + 1])


n_epi = 1
for _ in range(warm):
    body(n_epi, False); n_epi += 1
torch.cuda.synchronize()
count.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    body(n_epi, True); n_epi += 1
e1.record(); torch.cuda.synchronize()
env.check_status()
ms = e0.elapsed_time(e1)
print(which, kw, f"{int(count) / (ms / 1e3) / 1e6:.2f} M agent*steps/s, {ms / steps:.3f} ms/step (timed with per-phase syncs)",
      {k: round(v / steps, 3) for k, v in tot.items()}, "adam steps", [int(b._dev.adam_step) for b in brains])

"""Microbenchmark of the World kernels alone (random actions): CUDA-event time per launch + algorithmic GB/s."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reinlife_b200.World.vecworld import VecWorld

NW = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 30
target = int(sys.argv[3]) if len(sys.argv) > 3 else 100
iters = 30
vw = VecWorld(NW, H, W, 2, max_agents=target, seed=1)
vw.reset(); vw.top_up(target)
g = torch.Generator(device="cuda"); g.manual_seed(0)
ev = lambda: torch.cuda.Event(enable_timing=True)
ts = {"step": 0.0, "update": 0.0, "topup": 0.0}
agents = 0
for it in range(iters + 5):
    acts = torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g)
    vw.set_actions(acts)
    n = int(vw.n_agents.sum())
    e = [ev() for _ in range(4)]
    e[0].record(); vw.step(); e[1].record(); vw.update(); e[2].record(); vw.top_up(target); e[3].record()
    torch.cuda.synchronize()
    if it >= 5:
        agents += n
        ts["step"] += e[0].elapsed_time(e[1]); ts["update"] += e[1].elapsed_time(e[2]); ts["topup"] += e[2].elapsed_time(e[3])
C = H * W
navg = agents / iters / NW
out = {"n_worlds": NW, "grid": [H, W], "agents_per_world": navg}
for k in ts:
    ms = ts[k] / iters
    alg = NW * ((2 * C + 30 * navg if k == "step" else 0) + C + 12 * navg + 612 * navg)
    out[k] = {"ms": ms, "alg_GBs": alg / ms / 1e6}
print(json.dumps(out))

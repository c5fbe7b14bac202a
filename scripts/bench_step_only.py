"""Step kernel alone, repeated on a frozen world state (timing experiments; RL_WORLD_DEBUG variants)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reinlife_b200.World.vecworld import VecWorld
NW = 4096
vw = VecWorld(NW, 30, 30, 2, max_agents=100, seed=1)
vw.reset(); vw.top_up(100)
g = torch.Generator(device="cuda"); g.manual_seed(0)
tot = 0.0
for it in range(25):
    vw.set_actions(torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); vw.step(); e1.record(); torch.cuda.synchronize()
    if it >= 5: tot += e0.elapsed_time(e1)
    os.environ.pop("X", None)
    d = os.environ.pop("RL_WORLD_DEBUG", None)
    vw.update(); vw.top_up(100)
    if d: os.environ["RL_WORLD_DEBUG"] = d
print(os.environ.get("RL_WORLD_DEBUG", "0"), "step ms", tot / 20)

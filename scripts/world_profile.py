"""A few step / update+top-up launches at the bench shape (4096 x 30x30 x 100, float16 rows on) -- target for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reinlife_b200.World.vecworld import VecWorld
vw = VecWorld(4096, 30, 30, 2, max_agents=100, seed=1)
if os.environ.get("RL_NO_OBS16") is None:
    vw.enable_obs_fp16()
vw.reset(); vw.top_up(100)
g = torch.Generator(device="cuda"); g.manual_seed(0)
for it in range(int(os.environ.get("ITERS", 6))):
    vw.set_actions(torch.randint(0, 8, (4096, vw.S), device="cuda", dtype=torch.int8, generator=g))
    vw.step(); vw.update(top_up=100)
torch.cuda.synchronize()
print("done")

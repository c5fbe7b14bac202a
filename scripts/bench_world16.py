"""World kernel times with / without the float16 observation copies."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reinlife_b200.World.vecworld import VecWorld
for fp16 in (False, True, False, True):
    vw = VecWorld(4096, 30, 30, 2, max_agents=100, seed=1)
    if fp16:
        vw.enable_obs_fp16()
    vw.reset(); vw.top_up(100)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    ts = [0.0, 0.0]
    for it in range(25):
        vw.set_actions(torch.randint(0, 8, (4096, vw.S), device="cuda", dtype=torch.int8, generator=g))
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); vw.step(); e[1].record(); vw.update(top_up=100); e[2].record()
        torch.cuda.synchronize()
        if it >= 5:
            ts[0] += e[0].elapsed_time(e[1]) / 20; ts[1] += e[1].elapsed_time(e[2]) / 20
    print(f"fp16 copies {fp16}: step {ts[0]:.4f} ms, update+top-up {ts[1]:.4f} ms", flush=True)
    del vw

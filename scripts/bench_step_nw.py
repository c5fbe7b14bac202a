"""Step-kernel time vs number of worlds (wave quantisation experiment)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reinlife_b200.World.vecworld import VecWorld
for NW in [int(x) for x in sys.argv[1:]]:
    vw = VecWorld(NW, 30, 30, 2, max_agents=100, seed=1)
    vw.reset(); vw.top_up(100)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    tot = {"step": 0.0, "update": 0.0, "topup": 0.0}
    for it in range(25):
        vw.set_actions(torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g))
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); vw.step(); e[1].record(); vw.update(); e[2].record(); vw.top_up(100); e[3].record(); torch.cuda.synchronize()
        if it >= 5:
            for k, name in enumerate(tot): tot[name] += e[k].elapsed_time(e[k + 1])
    print(NW, {k: round(v / 20 * 1000, 1) for k, v in tot.items()}, "us; per 1024 worlds:", round(tot["step"] / 20 * 1000 / NW * 1024, 2))
    del vw
    torch.cuda.empty_cache()

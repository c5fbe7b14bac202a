"""One get_action pass of the dueling brains at the bench shape (4096 worlds x 100 agents) -- target for RL_TC_TRACE=1 / ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import reinlife_b200 as rl
from reinlife_b200.Models import PERD3QN
torch.manual_seed(0)
brains = [PERD3QN(exploration=0, capacity=64), PERD3QN(exploration=0, capacity=64)]
env = rl.Environment(width=30, height=30, brains=brains, max_agents=100, print_results=False, training=True,
                     n_worlds=int(os.environ.get("NW", 4096)), seed=0, device="cuda:0", precision="fp16")
env.reset(); env.top_up(100)
for n_epi in range(1, 4):
    env.act(n_epi); env.step(); env.update_env(n_epi, top_up=100)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for n_epi in range(4, 14):
    env.act(n_epi)
e1.record(); torch.cuda.synchronize()
print(f"act phase: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)

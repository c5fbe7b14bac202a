"""Small compute-sanitizer target for the batch-major tensor-core kernels (memcheck on the full-size target of
sanitize_events.py does not finish within 10 minutes on these spin-wait kernels): the paired event kernel on 1 / 21 events
(register-gather and TMA variants) and k_act_dueling_p on a 6-world batch.

    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_scale_gpu import _events_setup, _run_event_kernel, _half_copy      # noqa: E402
from brain_golden_util import state_dict                                     # noqa: E402

w0, tgt = state_dict("train_perd3qn/w0"), state_dict("train_perd3qn/target")
PART = os.environ.get("RL_SAN_PART", "all")          # "env": only the Environment steps (World kernels with float16 rows, TMA get_action, store, sampler)
for name, NW, per_world in (("1 event", 1, [1]), ("21 events", 7, [3, 1, 0, 5, 2, 4, 6])) if PART != "env" else ():
    z, vw, rows, rp, ring, n_ev, sidx = _events_setup(NW, per_world, seed=5)
    sd = torch.from_numpy(sidx).cuda()
    for ring_name, r in (("float32 ring", rp), ("float16 ring / TMA", _half_copy(rp))):
        g, l, p = _run_event_kernel("fp16p", vw, rows, r, w0, tgt, sd, n_ev)
        print(f"pair kernel, {name}, {ring_name}: grad[n] = {g[len(g) - 4]}, loss[0] = {l[0]:.4f}", flush=True)

import reinlife_b200 as rl                                                   # noqa: E402
from reinlife_b200.Models import PERD3QN                                     # noqa: E402
torch.manual_seed(0)
brains = [PERD3QN(exploration=0, train_freq=4, capacity=64), PERD3QN(exploration=0, train_freq=4, capacity=64)]
env = rl.Environment(width=12, height=12, brains=brains, max_agents=40, print_results=False, training=True,
                     n_worlds=6, seed=3, device="cuda:0", precision="fp16")
env.reset(); env.top_up(40)
for n_epi in range(1, 4):
    env.act(n_epi); env.step(); env.learn(n_epi); env.update_env(n_epi, top_up=40)
torch.cuda.synchronize()
print("Environment x3 steps (fp16, 6 worlds): adam steps", [int(b._dev.adam_step) for b in brains], flush=True)
print("sanitize small target done")

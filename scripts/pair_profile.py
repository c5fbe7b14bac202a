"""One launch of rl_brain_learn_p (and rl_brain_learn_h) at 20 480 events -- target for `ncu -k regex:k_learn_dueling`."""
import os, sys
import numpy as np
import torch
import ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_scale_gpu import _events_setup, _half_copy
from brain_golden_util import state_dict
from reinlife_b200 import _lib
from reinlife_b200.brains import DeviceBrain
w0, tgt = state_dict("train_perd3qn/w0"), state_dict("train_perd3qn/target")
z, vw, rows, rp, ring, n_ev, sidx = _events_setup(320, [64] * 320, seed=5)
if os.environ.get("RL_RING32") is None:
    rp = _half_copy(rp)      # the ring the Environment uses under precision="fp16" (TMA gather4 path)
brain = DeviceBrain(0, w0, "cuda"); brain.use_fp16 = True
brain.load_state_dict(tgt, target=True); brain.alloc_learn(rows.row_cap); brain.sample_idx[:n_ev] = torch.from_numpy(sidx).cuda()
st = vw._stream(); brain.build_wimg(st)
for fn in (vw.lib.rl_brain_learn_p, vw.lib.rl_brain_learn_h) if len(sys.argv) > 1 else (vw.lib.rl_brain_learn_p,):
    for _ in range(2):
        _lib.check(fn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs), C.c_void_p(brain.sample_idx.data_ptr()),
                      C.byref(brain.learn_bufs), C.c_void_p(brain.wimg_eh.data_ptr()), C.c_void_p(brain.wimg_th.data_ptr()), st))
    torch.cuda.synchronize()
print("done", n_ev)

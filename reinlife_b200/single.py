"""One-observation forward through the batched act kernel (the reference's per-agent `brain.get_action(state)` plugin
call, World/entities.py:215-222, Helpers/tester.py:58-68).  Slow by construction (one launch per agent); the
vectorised Environment.act() is the product path -- this exists so a brain object is usable on its own."""
import ctypes as C

import numpy as np
import torch

from . import _lib

_scratch = {}


def forward_single(brain, state):
    from .World.vecworld import VecWorld
    from .rows import RowLists
    from .brains import DeviceBrain
    dev = brain._dev.device if brain._dev is not None else torch.device("cuda", torch.cuda.current_device())
    key = str(dev)
    if key not in _scratch:
        vw = VecWorld(1, 3, 3, 1, max_agents=1, device=dev)
        rows = RowLists(vw)
        vw.n_agents[0] = 1
        rows.build(kinds_mask=1)
        _scratch[key] = (vw, rows, torch.zeros(1, dtype=torch.float64, device=dev),
                         torch.zeros((1, rows.row_cap, 8), device=dev))
    vw, rows, eps, q_out = _scratch[key]
    if brain._dev is None:
        brain._dev = DeviceBrain(brain.KIND, brain._host_sd, dev, lr=brain._lr(), gamma=brain._gamma(),
                                 batch=brain._batch(), has_target=brain.HAS_TARGET)
    state = np.asarray(state, np.float64).reshape(-1)
    if state.shape[0] != _lib.OBS_DIM:
        raise ValueError(f"expected a {_lib.OBS_DIM}-value observation, got {state.shape[0]}")
    row = torch.zeros(vw.ld, dtype=torch.float32)
    row[:_lib.OBS_DIM] = torch.from_numpy(state.astype(np.float32))
    vw.obs_state[0, 0].copy_(row.to(dev))
    acts = (_lib.BrainAct * 1)(_lib.BrainAct(brain._dev.kind, _lib.ACT_DQN, brain._dev.params.data_ptr(), eps.data_ptr()))
    with torch.cuda.device(dev):
        _lib.check(vw.lib.rl_brain_act_all(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), acts, 1, C.c_uint64(0),
                                           C.c_void_p(q_out.data_ptr()), None, vw._stream()))
    return q_out[0, 0].cpu().numpy()

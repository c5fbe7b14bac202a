"""Target for compute-sanitizer (memcheck / racecheck / synccheck) of the warp-specialised mbarrier kernels:
the fp16 and tf32 event kernels at >= 2 events per persistent CTA and the fp16 / tf32 get_action kernels at >= 2 tiles
per CTA, plus one full Environment step (world kernels, row lists, replay store / sample / priorities, Adam).

    compute-sanitizer --tool memcheck  python scripts/sanitize_events.py  > profiles/sanitizer_memcheck_r02.log
    compute-sanitizer --tool racecheck python scripts/sanitize_events.py  > profiles/sanitizer_racecheck_r02.log
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def events(n_per_cta=3):
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, ReplayRings
    from reinlife_b200.Models import packing
    from test_learn_gpu import _mk, _fake_events
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    NW = 32
    vw, rows = _mk(NW)
    grid = vw.lib.rl_learn_grid()
    per_world = [(n_per_cta * grid + NW - 1) // NW] * NW
    rp = ReplayRings(NW, 128, "cuda")
    rp.obs.copy_(torch.rand_like(rp.obs)); rp.next_obs.copy_(torch.rand_like(rp.next_obs))
    rp.obs[..., 153:] = 0; rp.next_obs[..., 153:] = 0
    rp.action.copy_(torch.randint(0, 8, rp.action.shape, dtype=torch.int8, device="cuda"))
    rp.reward.copy_(torch.randn_like(rp.reward)); rp.len[:] = 128
    n_ev = _fake_events(vw, rows, per_world)
    sidx = torch.from_numpy(rng.integers(0, 128, size=(n_ev, 64)).astype(np.int32)).cuda()
    for mode in ("fp16p", "fp16", "tf32", "fp32"):
        brain = DeviceBrain(0, packing.default_init(0), "cuda")
        brain.use_fp16 = mode in ("fp16", "fp16p")
        brain.alloc_learn(rows.row_cap)
        brain.sample_idx[:n_ev] = sidx
        st = vw._stream()
        if mode == "fp32":
            _lib.check(vw.lib.rl_brain_learn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                             C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs), st))
        else:
            brain.build_wimg(st)
            fn = {"fp16": vw.lib.rl_brain_learn_h, "fp16p": vw.lib.rl_brain_learn_p}.get(mode, vw.lib.rl_brain_learn_tc)
            we, wt = (brain.wimg_eh, brain.wimg_th) if mode in ("fp16", "fp16p") else (brain.wimg_e, brain.wimg_t)
            _lib.check(fn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs), C.c_void_p(brain.sample_idx.data_ptr()),
                          C.byref(brain.learn_bufs), C.c_void_p(we.data_ptr()), C.c_void_p(wt.data_ptr()), st))
        torch.cuda.synchronize()
        print(f"event kernel {mode}: {n_ev} events over {grid} CTAs, grad[n]={float(brain.grad[brain.dims.n_train])}", flush=True)
    # the paired kernel's TMA variant: float16 ring -> tile::gather4 row gathers, accumulator warps (setmaxnreg)
    r16 = ReplayRings(NW, 128, "cuda", fp16=True)
    r16.obs.copy_(rp.obs.half()); r16.next_obs.copy_(rp.next_obs.half())
    r16.obs[..., -1] = 1.0; r16.next_obs[..., -1] = 1.0
    for name in ("action", "reward", "done", "prio", "pw", "len", "pos"):
        getattr(r16, name).copy_(getattr(rp, name))
    brain = DeviceBrain(0, packing.default_init(0), "cuda")
    brain.use_fp16 = True
    brain.alloc_learn(rows.row_cap)
    brain.sample_idx[:n_ev] = sidx
    st = vw._stream()
    brain.build_wimg(st)
    _lib.check(vw.lib.rl_brain_learn_p(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(r16.bufs), C.c_void_p(brain.sample_idx.data_ptr()),
                                       C.byref(brain.learn_bufs), C.c_void_p(brain.wimg_eh.data_ptr()), C.c_void_p(brain.wimg_th.data_ptr()), st))
    torch.cuda.synchronize()
    print(f"event kernel fp16p, float16 ring (TMA gather4): {n_ev} events, grad[n]={float(brain.grad[brain.dims.n_train])}", flush=True)


def env_steps():
    import reinlife_b200 as rl
    from reinlife_b200.Models import PERD3QN
    torch.manual_seed(0)
    for precision in ("fp16", "tf32"):
        brains = [PERD3QN(exploration=0, train_freq=4, capacity=256), PERD3QN(exploration=0, train_freq=4, capacity=256)]
        env = rl.Environment(width=30, height=30, brains=brains, max_agents=100, print_results=False, training=True,
                             n_worlds=384, seed=3, device="cuda:0", precision=precision)
        env.reset(); env.top_up(100)
        for n_epi in range(1, 4):
            env.act(n_epi); env.step(); env.learn(n_epi); env.update_env(n_epi, top_up=100)
        torch.cuda.synchronize()
        print(f"Environment x3 steps ({precision}): adam steps", [int(b._dev.adam_step) for b in brains], flush=True)


if __name__ == "__main__":
    part = os.environ.get("RL_SAN_PART", "all")
    if part in ("all", "events"):
        events()
    if part in ("all", "env"):
        env_steps()
    print("sanitize target done")

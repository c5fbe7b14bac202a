"""Quick check + timing of the paired event kernel (rl_brain_learn_p) against the fp32 kernel and rl_brain_learn_h."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_scale_gpu import _events_setup, _run_event_kernel, _check_tc_vs, _half_copy
from brain_golden_util import state_dict
from reinlife_b200.Models import packing
import ctypes as C
from reinlife_b200 import _lib
from reinlife_b200.brains import DeviceBrain

w0, tgt = state_dict("train_perd3qn/w0"), state_dict("train_perd3qn/target")
d, m = packing.dims(0), packing.grad_mask(0)
cases = [("21 events", 7, [3, 1, 0, 5, 2, 4, 6]), ("1 event", 1, [1]), ("613 events", 64, None), ("20k events", 320, [64] * 320)]
for name, NW, per_world in cases:
    if per_world is None:
        per_world = np.random.default_rng(11).integers(5, 15, NW).tolist()
    z, vw, rows, rp, ring, n_ev, sidx = _events_setup(NW, per_world, seed=5)
    sd = torch.from_numpy(sidx).cuda()
    ref = _run_event_kernel("fp32", vw, rows, rp, w0, tgt, sd, n_ev)
    r16 = _half_copy(rp)
    for ring_name, ring_used in (("float32 ring (register gather)", rp), ("float16 ring (TMA gather4)", r16)):
        got = _run_event_kernel("fp16p", vw, rows, ring_used, w0, tgt, sd, n_ev)
        try:
            _check_tc_vs(ref, got, m, d, n_ev, name)
            print(name, ring_name, "OK", flush=True)
        except AssertionError as e:
            print(name, ring_name, "FAIL", str(e)[:600], flush=True)
    if n_ev > 10000:
        for mode, rp in (("fp16", rp), ("fp16p", rp), ("fp16p", r16)):
            brain = DeviceBrain(0, w0, "cuda"); brain.use_fp16 = True
            brain.load_state_dict(tgt, target=True); brain.alloc_learn(rows.row_cap); brain.sample_idx[:n_ev] = sd
            st = vw._stream(); brain.build_wimg(st)
            fn = vw.lib.rl_brain_learn_p if mode == "fp16p" else vw.lib.rl_brain_learn_h
            def run():
                _lib.check(fn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs), C.c_void_p(brain.sample_idx.data_ptr()),
                              C.byref(brain.learn_bufs), C.c_void_p(brain.wimg_eh.data_ptr()), C.c_void_p(brain.wimg_th.data_ptr()), st))
            for _ in range(3): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"{mode} ({'float16' if rp.fp16 else 'float32'} ring): {ms:.3f} ms per launch for {n_ev} events -> {n_ev * 24.89e6 / ms / 1e9:.1f} TFLOP/s", flush=True)

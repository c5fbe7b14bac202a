"""GPU parity of the persistent kernels in the regime the benchmark runs them in: every CTA of rl_learn_grid() walks
SEVERAL events / tiles (the parity tests of test_learn_gpu.py / test_tc_gpu.py stay below one event per CTA).

* tensor-core event kernels (k_learn_dueling_h = fp16 operands, the default; k_learn_dueling_tc2 = tf32) on >= 600
  events (>= 4 per CTA) against (a) the oracle's per-event train() restatement directly and (b) the fp32 kernel, and on
  >= 20 000 events (one bench step's worth, ~140 per CTA) against the fp32 kernel + the oracle on sampled events.
  What only shows with several events per CTA: dW2 resident in TMEM across events, the weight-chunk ring wrap, the
  two-phase next-event metadata, next-event row prefetch, the go/done mbarrier phase flips.
* tensor-core get_action kernels on >= 600 tiles per brain against the fp32 kernel and the oracle forward.
* the fp32 SIMT tile kernels (k_learn_dueling, k_learn_dqn, k_ppo_tiles, k_brain_act) with >= 3 tiles per CTA against
  the oracle.
Tolerances (stated): fp32 kernels -- the ones of test_learn_gpu.py; tensor-core kernels (11 significant bits per operand,
fp32 accumulation) -- summed gradients within 1 % of each tensor's gradient scale, per-event loss within 2 %,
priorities within 2 % (+ 2e-2 absolute), Q within 2 % of the Q scale, >= 99 % identical greedy actions.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from brain_golden_util import golden, state_dict          # noqa: E402
from test_learn_gpu import _mk, _fake_events, _golden2, _sd2, _ppo_manual_plan   # noqa: E402

pytestmark = pytest.mark.gpu

PARTS = (("W1", "off_w1t", "off_b1"), ("b1", "off_b1", "off_w2t"), ("W2", "off_w2t", "off_b2"), ("b2", "off_b2", "off_wh"),
         ("Wh", "off_wh", "off_bh"))


def _fill_rings(rp, rng, obs_all, n_items):
    """Every ring gets n_items transitions drawn from the golden observation pool (one upload per array)."""
    NW = rp.n_worlds
    pool = torch.from_numpy(np.pad(obs_all.astype(np.float32), ((0, 0), (0, 7)))).cuda()
    io = torch.from_numpy(rng.integers(0, len(obs_all), (NW, n_items))).cuda()
    jo = torch.from_numpy(rng.integers(0, len(obs_all), (NW, n_items))).cuda()
    rp.obs[:, :n_items] = pool[io]
    rp.next_obs[:, :n_items] = pool[jo]
    a = rng.integers(0, 8, (NW, n_items)).astype(np.int8)
    r = rng.choice(np.array([0.0, 0.2, 0.5, -3.0, -20.0], np.float32), (NW, n_items))
    d = (r < 0).astype(np.uint8)
    rp.action[:, :n_items] = torch.from_numpy(a).cuda()
    rp.reward[:, :n_items] = torch.from_numpy(r).cuda()
    rp.done[:, :n_items] = torch.from_numpy(d).cuda()
    rp.len[:] = n_items
    return io.cpu().numpy(), jo.cpu().numpy(), a, r, d


def _run_event_kernel(mode, vw, rows, rp, w0, tgt, sidx, n_ev):
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain
    brain = DeviceBrain(0, w0, "cuda", lr=1e-3, gamma=0.99)
    brain.use_fp16 = mode in ("fp16", "fp16p")
    brain.load_state_dict(tgt, target=True)
    brain.alloc_learn(rows.row_cap)
    brain.sample_idx[:n_ev] = sidx
    st = vw._stream()
    if mode == "fp32":
        _lib.check(vw.lib.rl_brain_learn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                         C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs), st))
    elif mode == "tf32":
        brain.build_wimg(st)
        _lib.check(vw.lib.rl_brain_learn_tc(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                            C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs),
                                            C.c_void_p(brain.wimg_e.data_ptr()), C.c_void_p(brain.wimg_t.data_ptr()), st))
    else:
        brain.build_wimg(st)
        fn = vw.lib.rl_brain_learn_p if mode == "fp16p" else vw.lib.rl_brain_learn_h     # fp16p: two events per CTA iteration
        _lib.check(fn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                           C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs),
                                           C.c_void_p(brain.wimg_eh.data_ptr()), C.c_void_p(brain.wimg_th.data_ptr()), st))
    torch.cuda.synchronize()
    return (brain.grad.cpu().numpy().copy(), brain.loss[:n_ev].cpu().numpy().copy(),
            brain.new_prio[:n_ev].cpu().numpy().copy())


def _grad_parts(g, d):
    out = {}
    for name, lo, hi in PARTS:
        lo = 0 if lo == "off_w1t" else getattr(d, lo)
        out[name] = (lo, getattr(d, hi))
    out["bh"] = (d.off_bh, d.off_bh + 9)
    return out


def _check_tc_vs(ref, got, mask, d, n_ev, what, grad_tol=1e-2):
    g32, l32, p32 = ref
    gtc, ltc, ptc = got
    nt = len(g32) - 4
    assert gtc[nt] == n_ev, (what, gtc[nt], n_ev)
    for name, (lo, hi) in _grad_parts(g32, d).items():
        a, b = g32[lo:hi] * mask[lo:hi], gtc[lo:hi] * mask[lo:hi]
        scale = np.abs(a).max()
        err = np.abs(a - b).max() / scale
        assert err < grad_tol, (what, name, err, scale)
    np.testing.assert_allclose(ltc, l32, rtol=2e-2, atol=1e-3, err_msg=what)
    np.testing.assert_allclose(ptc, p32, rtol=2e-2, atol=2e-2, err_msg=what)


def _events_setup(NW, per_world, seed, n_items=200, cap=256):
    from reinlife_b200.brains import ReplayRings
    z = golden()
    rng = np.random.default_rng(seed)
    vw, rows = _mk(NW)
    rp = ReplayRings(NW, cap, "cuda")
    ring = _fill_rings(rp, rng, z["obs"], n_items)
    n_ev = _fake_events(vw, rows, per_world)
    sidx = rng.integers(0, n_items, size=(n_ev, 64)).astype(np.int32)
    return z, vw, rows, rp, ring, n_ev, sidx


def _oracle_event(z, ring, w, idx):
    io, jo, a, r, d = ring
    return (z["obs"][io[w, idx]], a[w, idx].astype(np.int64), r[w, idx].astype(np.float64), z["obs"][jo[w, idx]],
            d[w, idx].astype(np.float64))


def test_event_kernels_600_events_vs_oracle_and_fp32():
    """>= 4 events per persistent CTA.  The oracle (explicit fp32 restatement of PERD3QNAgent.train(), PERD3QN.py:94-115,
    pinned to the reference) is the checker for ALL three kernels here -- the tensor-core kernels are compared with it
    directly, not only with another CUDA kernel."""
    from reinlife_b200.Models import packing
    from oracle import brain_oracle as bo
    NW = 64
    rngp = np.random.default_rng(11)
    per_world = rngp.integers(5, 15, NW).tolist()
    per_world[7] = 0
    z, vw, rows, rp, ring, n_ev, sidx = _events_setup(NW, per_world, seed=5)
    assert n_ev >= 600 and n_ev >= 4 * vw.lib.rl_learn_grid()
    w0, tgt = state_dict("train_perd3qn/w0"), state_dict("train_perd3qn/target")
    sidx_d = torch.from_numpy(sidx).cuda()
    out = {m: _run_event_kernel(m, vw, rows, rp, w0, tgt, sidx_d, n_ev) for m in ("fp32", "tf32", "fp16", "fp16p")}
    events, e = [], 0
    for w in range(NW):
        for _ in range(per_world[w]):
            events.append(_oracle_event(z, ring, w, sidx[e])); e += 1
    g_ref, losses, prios = bo.dueling_batched_update(w0, tgt, events, 0.99)
    d, m = packing.dims(0), packing.grad_mask(0)
    nt = d.n_train
    # (a) fp32 kernel vs oracle at the fp32 tolerance of test_learn_gpu.py
    g32 = out["fp32"][0]
    assert g32[nt] == n_ev
    got = packing.unpack(0, np.concatenate([g32[:nt] / n_ev * m, np.zeros(d.n_total - nt, np.float32)]))
    for k in g_ref:
        scale = max(1.0, np.abs(g_ref[k]).max())
        np.testing.assert_allclose(got[k].numpy(), g_ref[k], rtol=2e-4, atol=2e-5 * scale, err_msg=k)
    np.testing.assert_allclose(out["fp32"][1], np.array(losses), rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(out["fp32"][2], np.stack(prios), rtol=2e-4, atol=2e-4)
    # (b) tensor-core kernels vs the ORACLE directly: gradients (packed into the kernel layout), per-event loss, priorities
    flat_ref = packing.pack(0, {k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in g_ref.items()})[:nt] * n_ev
    ref = (np.concatenate([flat_ref, np.zeros(4, np.float32)]), np.array(losses, np.float32), np.stack(prios).astype(np.float32))
    for mode in ("tf32", "fp16", "fp16p"):
        _check_tc_vs(ref, out[mode], m, d, n_ev, f"{mode} vs oracle")
        _check_tc_vs(out["fp32"], out[mode], m, d, n_ev, f"{mode} vs fp32 kernel")


def test_event_kernels_bench_scale_20k_events():
    """One bench step's worth of events for one brain (>= 20 000, ~140 per persistent CTA): tensor-core kernels vs the
    fp32 kernel on everything, and all three vs the oracle on 48 sampled events (first / last waves of every CTA
    stride included)."""
    from reinlife_b200.Models import packing
    from oracle import brain_oracle as bo
    NW = 320
    per_world = [64] * NW
    per_world[3] = 0; per_world[100] = 17
    z, vw, rows, rp, ring, n_ev, sidx = _events_setup(NW, per_world, seed=8)
    assert n_ev >= 20000
    w0, tgt = state_dict("train_perd3qn/w0"), state_dict("train_perd3qn/target")
    sidx_d = torch.from_numpy(sidx).cuda()
    out = {m: _run_event_kernel(m, vw, rows, rp, w0, tgt, sidx_d, n_ev) for m in ("fp32", "tf32", "fp16", "fp16p")}
    d, m = packing.dims(0), packing.grad_mask(0)
    for mode in ("tf32", "fp16", "fp16p"):
        _check_tc_vs(out["fp32"], out[mode], m, d, n_ev, f"{mode} vs fp32 kernel, {n_ev} events")
    ev_world = np.repeat(np.arange(NW), per_world)
    rng = np.random.default_rng(2)
    picks = sorted(set(rng.integers(0, n_ev, 40).tolist() + [0, 1, 147, 148, 149, n_ev - 149, n_ev - 2, n_ev - 1]))
    for e in picks:
        _, loss, prio = bo.dueling_event_grads(w0, tgt, *_oracle_event(z, ring, int(ev_world[e]), sidx[e]), 0.99)
        np.testing.assert_allclose(out["fp32"][1][e], loss, rtol=2e-4, atol=1e-4, err_msg=f"fp32 event {e}")
        np.testing.assert_allclose(out["fp32"][2][e], prio, rtol=2e-4, atol=2e-4, err_msg=f"fp32 event {e}")
        for mode in ("tf32", "fp16", "fp16p"):
            np.testing.assert_allclose(out[mode][1][e], loss, rtol=2e-2, atol=1e-3, err_msg=f"{mode} event {e}")
            np.testing.assert_allclose(out[mode][2][e], prio, rtol=2e-2, atol=2e-2, err_msg=f"{mode} event {e}")


def test_event_kernels_are_deterministic_run_to_run():
    """Same inputs twice -> bit-identical loss / priorities; gradients identical up to the order of the red.global.add
    flushes (fp16 / tf32 kernels) -- the fp32 kernel's fixed-order slab reduction is bit-identical."""
    NW = 40
    per_world = [12] * NW
    z, vw, rows, rp, ring, n_ev, sidx = _events_setup(NW, per_world, seed=21)
    w0, tgt = state_dict("train_perd3qn/w0"), state_dict("train_perd3qn/target")
    sidx_d = torch.from_numpy(sidx).cuda()
    for mode in ("fp32", "fp16", "fp16p", "tf32"):
        a = _run_event_kernel(mode, vw, rows, rp, w0, tgt, sidx_d, n_ev)
        b = _run_event_kernel(mode, vw, rows, rp, w0, tgt, sidx_d, n_ev)
        assert (a[1] == b[1]).all() and (a[2] == b[2]).all(), mode
        if mode == "fp32":
            assert (a[0] == b[0]).all()
        else:
            assert np.abs(a[0] - b[0]).max() <= 1e-4 * np.abs(a[0]).max(), mode


def test_act_kernels_600_tiles_per_brain():
    """get_action over >= 600 64-row tiles per brain (>= 4 per persistent CTA): tensor-core forwards (tf32, fp16) vs the
    fp32 kernel on every row, and all three vs the oracle forward (PERD3QN.py:198-202 at B = 1) on 2 000 sampled rows."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain
    from reinlife_b200.World.vecworld import VecWorld
    from reinlife_b200.rows import RowLists
    from oracle import brain_oracle as bo
    NW = 800
    vw = VecWorld(NW, 30, 30, 2, max_agents=100, seed=12)
    rows = RowLists(vw, row_cap=NW * 128)
    vw.reset(); vw.top_up(100)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for _ in range(2):
        vw.set_actions(torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g))
        vw.step(); vw.update(); vw.top_up(100)
    rows.build(kinds_mask=1)
    sds = [state_dict("perd3qn"), state_dict("d3qn")]
    brains = [DeviceBrain(0, sd, "cuda") for sd in sds]
    eps = torch.tensor([0.3, 0.0], dtype=torch.float64, device="cuda")
    descs = (_lib.BrainAct * 2)(*[b.act_desc(_lib.ACT_DUELING, eps.data_ptr() + 8 * i) for i, b in enumerate(brains)])
    out = {}
    for mode in ("fp32", "tf32", "fp16", "fp16p"):       # fp16p = rl_brain_act_p (batch-major 128-row tiles, the default)
        q = torch.zeros((2, rows.row_cap, 8), device="cuda")
        vw.rec[:, :, 13] = 255
        if mode == "fp32":
            _lib.check(vw.lib.rl_brain_act_all(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), descs, 2, C.c_uint64(7),
                                               C.c_void_p(q.data_ptr()), None, vw._stream()))
        else:
            for i, b in enumerate(brains):
                b.use_fp16 = mode != "tf32"
                b.build_wimg(vw._stream())
                fn = vw.lib.rl_brain_act_tc if mode == "tf32" else vw.lib.rl_brain_act_p if mode == "fp16p" else vw.lib.rl_brain_act_h
                img = b.wimg_e if mode == "tf32" else b.wimg_eh
                _lib.check(fn(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), i, C.byref(descs[i]),
                              C.c_void_p(img.data_ptr()), C.c_uint64(7), C.c_void_p(q.data_ptr()), vw._stream()))
        torch.cuda.synchronize()
        out[mode] = (q.cpu().numpy(), vw.rec[:, :, 13].cpu().numpy().view(np.int8).copy())
    a32 = out["fp32"][1]
    listed = a32 != -1
    obs = vw.obs_state.view(-1, vw.ld)
    rng = np.random.default_rng(4)
    for i in range(2):
        n = int(rows.total[i * 3])
        assert n >= 600 * 64, n
        ids = rows.rows[i * 3, :n].cpu().numpy()
        pick = np.unique(np.concatenate([rng.integers(0, n, 2000), np.arange(64), np.arange(n - 64, n)]))
        x = obs[torch.from_numpy(ids[pick]).cuda().long(), :153].cpu().numpy()
        q_or = bo.dueling_forward(sds[i], x, per_row_mean=True)
        q32 = out["fp32"][0][i, :n]
        np.testing.assert_allclose(q32[pick], q_or, rtol=1e-4, atol=1e-4)
        scale = np.abs(q32).max()
        for mode in ("tf32", "fp16", "fp16p"):
            qtc = out[mode][0][i, :n]
            assert np.abs(qtc[pick] - q_or).max() < 2e-2 * scale, (mode, "vs oracle")
            assert np.abs(q32 - qtc).max() < 2e-2 * scale, mode
            assert (q32.argmax(1) == qtc.argmax(1)).mean() >= 0.99, mode
    for mode in ("tf32", "fp16", "fp16p"):
        atc = out[mode][1]
        assert ((atc != -1) == listed).all(), mode
        assert (a32[listed] == atc[listed]).mean() >= 0.99, mode


def test_dqn_tile_kernel_many_tiles_vs_oracle():
    """k_learn_dqn with >= 3 tiles per persistent CTA (two 32-row events per tile): mean of per-event gradients and
    per-event smooth-L1 losses vs the oracle (Models/DQN.py:142-153), skipped events (-1) not counted."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, ReplayRings
    from reinlife_b200.Models import packing
    from oracle import brain_oracle as bo
    z, z1 = _golden2(), golden()
    rng = np.random.default_rng(13)
    NW, cap, n_items = 64, 128, 100
    per_world = [16] * NW
    vw, rows = _mk(NW)
    w0, tgt = _sd2(z, "train_dqn/w0"), _sd2(z, "train_dqn/target")
    brain = DeviceBrain(1, w0, "cuda", lr=5e-4, gamma=0.98, batch=32)
    brain.load_state_dict(tgt, target=True)
    brain.alloc_learn(rows.row_cap)
    rp = ReplayRings(NW, cap, "cuda", prioritized=False)
    io, jo, a, r, d = _fill_rings(rp, rng, z1["obs"], n_items)
    n_ev = _fake_events(vw, rows, per_world)
    assert n_ev >= 2 * 3 * vw.lib.rl_learn_grid()
    sidx = rng.integers(0, n_items, size=(n_ev, 32)).astype(np.int32)
    skipped = set(rng.integers(0, n_ev, 25).tolist())
    for e in skipped:
        sidx[e] = -1
    brain.sample_idx[:n_ev] = torch.from_numpy(sidx).cuda()
    _lib.check(vw.lib.rl_brain_learn_dqn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                         C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs), vw._stream()))
    torch.cuda.synchronize()
    acc, losses = None, {}
    ev_world = np.repeat(np.arange(NW), per_world)
    obs = z1["obs"]
    for e in range(n_ev):
        if e in skipped:
            continue
        w, i = int(ev_world[e]), sidx[e]
        g, loss = bo.dqn_iter_grads(w0, tgt, obs[io[w, i]], a[w, i].astype(np.int64), r[w, i].astype(np.float64), obs[jo[w, i]],
                                    1.0 - d[w, i].astype(np.float64))
        losses[e] = loss
        acc = g if acc is None else {k: acc[k] + g[k] for k in g}
    n_valid = n_ev - len(skipped)
    grad = brain.grad.cpu().numpy()
    nt = brain.dims.n_train
    assert grad[nt] == n_valid
    got = packing.unpack(1, np.concatenate([grad[:nt] / n_valid * packing.grad_mask(1), np.zeros(brain.dims.n_total - nt, np.float32)]))
    for k in acc:
        ref = acc[k] / n_valid
        np.testing.assert_allclose(got[k].numpy(), ref, rtol=5e-4, atol=2e-5 * max(1.0, np.abs(ref).max()), err_msg=k)
    loss_dev = brain.loss[:n_ev].cpu().numpy()
    for e, l in losses.items():
        np.testing.assert_allclose(loss_dev[e], l, rtol=2e-4, atol=1e-4)


def test_ppo_tile_kernels_many_tiles_vs_oracle():
    """k_ppo_tiles<0/1> + k_ppo_gae over >= 3 tiles per persistent CTA (ragged segments of 1..150 rows that straddle
    tiles): mean over segments of the per-segment gradient vs the oracle (Models/PPO.py:136-162)."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, PpoData
    from reinlife_b200.Models import packing
    from oracle import brain_oracle as bo
    z, z1 = _golden2(), golden()
    rng = np.random.default_rng(17)
    w0 = _sd2(z, "train_ppo/w0")
    obs_all = z1["obs"]
    NW = 60
    vw, rows = _mk(NW)
    segs, n_rows = [], 0
    for w in range(NW):
        cur, left = [], 500
        while left > 0:
            T = int(min(left, rng.choice([1, 2, 7, 20, 64, 65, 150])))
            o = obs_all[rng.integers(0, 512, T)]; no = obs_all[rng.integers(0, 512, T)]
            pi, _ = bo.ppo_forward(w0, o)
            a = rng.integers(0, 8, T)
            cur.append(dict(obs=o, next_obs=no, action=a, reward=rng.choice([0.0, 0.002, 0.005, -0.03, -0.42], T),
                            prob_a=(pi[np.arange(T), a] * rng.uniform(0.6, 1.4, T)).astype(np.float32), done=rng.random(T) < 0.2))
            left -= T; n_rows += T
        segs.append(cur)
    assert n_rows >= 3 * 64 * vw.lib.rl_learn_grid()
    brain = DeviceBrain(2, w0, "cuda", lr=5e-4, gamma=0.98, batch=64, has_target=False)
    brain.alloc_learn(rows.row_cap, need_batch_bufs=False)
    pd = PpoData(NW, 512, NW * 512, "cuda")
    n_ev = _ppo_manual_plan(pd, rows, segs)
    _lib.check(vw.lib.rl_ppo_epoch(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(pd.bufs), C.byref(brain.learn_bufs), vw._stream()))
    torch.cuda.synchronize()
    assert int(pd.status) == 0
    acc = None
    for ss in segs:
        for sg in ss:
            g, _ = bo.ppo_epoch_grads(w0, sg["obs"], sg["action"], sg["reward"], sg["next_obs"], sg["prob_a"], sg["done"])
            acc = g if acc is None else {k: acc[k] + g[k] for k in g}
    grad = brain.grad.cpu().numpy()
    nt = brain.dims.n_train
    assert grad[nt] == n_ev
    got = packing.unpack(2, np.concatenate([grad[:nt] / n_ev * packing.grad_mask(2), np.zeros(brain.dims.n_total - nt, np.float32)]))
    for k in acc:
        ref = acc[k] / n_ev
        np.testing.assert_allclose(got[k].numpy(), ref, rtol=1e-3, atol=2e-6 + 5e-5 * np.abs(ref).max(), err_msg=k)


def _half_copy(rp):
    """float16 twin of a float32 ring (what rl_replay_store writes when obs_fp16 = 1: rounded rows, last column = 1)."""
    from reinlife_b200.brains import ReplayRings
    r16 = ReplayRings(rp.n_worlds, rp.capacity, "cuda", prioritized=rp.prioritized, fp16=True)
    r16.obs.copy_(rp.obs.half()); r16.next_obs.copy_(rp.next_obs.half())
    r16.obs[..., -1] = 1.0; r16.next_obs[..., -1] = 1.0
    for name in ("action", "reward", "done", "prio", "pw", "len", "pos"):
        getattr(r16, name).copy_(getattr(rp, name))
    return r16


def test_float16_ring_pair_kernel_and_fp32_kernel():
    """precision="fp16" keeps the dueling brains' replay rows as float16 (rl_replay_bufs.obs_fp16).  The paired event kernel
    must give the SAME results from the float16 ring as from the float32 ring (it rounds the same rows to the same halves
    at gather time), and the fp32 kernel reading the float16 ring (bench.py's parity_check) stays within the tensor-core
    tolerance of the fp32 kernel on the float32 ring."""
    from reinlife_b200.Models import packing
    NW = 48
    per_world = [13] * NW
    per_world[5] = 0
    z, vw, rows, rp, ring, n_ev, sidx = _events_setup(NW, per_world, seed=31)
    r16 = _half_copy(rp)
    w0, tgt = state_dict("train_perd3qn/w0"), state_dict("train_perd3qn/target")
    sd = torch.from_numpy(sidx).cuda()
    ref32 = _run_event_kernel("fp32", vw, rows, rp, w0, tgt, sd, n_ev)
    p32 = _run_event_kernel("fp16p", vw, rows, rp, w0, tgt, sd, n_ev)
    p16 = _run_event_kernel("fp16p", vw, rows, r16, w0, tgt, sd, n_ev)
    f16 = _run_event_kernel("fp32", vw, rows, r16, w0, tgt, sd, n_ev)
    assert (p32[1] == p16[1]).all() and (p32[2] == p16[2]).all()                  # per-event loss / priorities: identical
    assert np.abs(p32[0] - p16[0]).max() <= 1e-4 * np.abs(p32[0]).max()           # gradients: red.add order only
    d, m = packing.dims(0), packing.grad_mask(0)
    _check_tc_vs(ref32, p16, m, d, n_ev, "pair kernel, float16 ring vs fp32 kernel, float32 ring")
    _check_tc_vs(ref32, f16, m, d, n_ev, "fp32 kernel, float16 ring vs float32 ring")
    _check_tc_vs(f16, p16, m, d, n_ev, "pair kernel vs fp32 kernel, both on the float16 ring")


def test_replay_store_writes_float16_rows():
    """rl_replay_store into a float16 ring == float16(row) of what it writes into a float32 ring, last column = 1."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import ReplayRings
    from reinlife_b200.World.vecworld import VecWorld
    from reinlife_b200.rows import RowLists
    NW, cap = 5, 64
    vw = VecWorld(NW, 12, 12, 2, max_agents=40, seed=3)
    rows = RowLists(vw)
    vw.enable_obs_fp16()                      # ring c: copied from the float16 rows the World kernels emit
    hs, hp = vw.obs_state_h.data_ptr(), vw.obs_prime_h.data_ptr()
    vw.reset(); vw.top_up(40)
    a, b, c = ReplayRings(NW, cap, "cuda"), ReplayRings(NW, cap, "cuda", fp16=True), ReplayRings(NW, cap, "cuda", fp16=True)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    for _ in range(4):
        vw.set_actions(torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g))
        vw.step()
        rows.build(kinds_mask=6, train_freq=[3, 3], event_on=[1, 1])
        for rp in (a, b, c):
            vw.bufs.obs_state_h, vw.bufs.obs_prime_h = (hs, hp) if rp is c else (None, None)
            _lib.check(vw.lib.rl_replay_store(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), 0, C.byref(rp.bufs), vw._stream()))
        vw.update(); vw.top_up(40)
    torch.cuda.synchronize()
    assert (a.len == b.len).all() and (a.pos == b.pos).all() and int(a.len.max()) == cap
    for x, y, z in ((a.obs, b.obs, c.obs), (a.next_obs, b.next_obs, c.next_obs)):
        want = x.half().clone(); want[..., -1] = 1.0
        filled = torch.arange(cap, device="cuda")[None, :] < a.len[:, None]
        assert torch.equal(y[filled], want[filled]) and torch.equal(z[filled], want[filled])
    for r in (b, c):
        assert torch.equal(a.prio, r.prio) and torch.equal(a.action, r.action) and torch.equal(a.reward, r.reward)


def test_replay_max_priority_is_maintained_exactly():
    """rl_replay_store no longer scans the priority array for max(priorities) (PERD3QN.py:147): the store / priority-update
    kernels keep {max, count of entries holding it} per ring (rl_replay_bufs.maxst).  Over rounds of stores (ring wrap) and
    priority updates with duplicate indices, values above / equal to / below the maximum and updates that remove every
    holder of the maximum, the state equals a recount of the array whenever it is known, and every new item is stored with
    the true maximum."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import ReplayRings
    from reinlife_b200.World.vecworld import VecWorld
    from reinlife_b200.rows import RowLists
    NW, cap, B = 6, 96, 64
    vw = VecWorld(NW, 12, 12, 2, max_agents=40, seed=3)
    rows = RowLists(vw)
    vw.reset(); vw.top_up(40)
    rp = ReplayRings(NW, cap, "cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    rng = np.random.default_rng(1)
    known = 0
    for step in range(30):
        vw.set_actions(torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g))
        vw.step()
        rows.build(kinds_mask=6, train_freq=[3, 3], event_on=[1, 1])
        before = rp.prio.clone(); len0 = rp.len.clone(); pos0 = rp.pos.clone()
        _lib.check(vw.lib.rl_replay_store(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), 0, C.byref(rp.bufs), vw._stream()))
        torch.cuda.synchronize()
        cnt = rows.count[_lib.ROWS_STORE].cpu().numpy()
        for w in range(NW):
            if cnt[w] == 0:
                continue
            want = float(before[w].max()) if int(len0[w]) > 0 else 1.0
            slots = [(int(pos0[w]) + k) % cap for k in range(max(0, cnt[w] - cap), cnt[w])]
            assert (rp.prio[w, slots].cpu().numpy() == np.float32(want)).all(), (step, w)
        st = rp.maxst.cpu().numpy()
        for w in range(NW):
            if st[w, 1] > 0:
                known += 1
                m = float(rp.prio[w].max())
                assert np.int32(st[w, 0]).view(np.float32) == np.float32(m) and st[w, 1] == int((rp.prio[w] == m).sum()), (step, w, "store")
        # priority update on hand-made events: duplicates inside and across events, values above / equal / below the maximum
        n_ev = rng.integers(0, 4, NW)
        ev_tot = _fake_events_scale(vw, rows, n_ev)
        if ev_tot == 0:
            vw.update(); vw.top_up(40); continue
        L = np.maximum(rp.len.cpu().numpy(), 1)
        sidx = np.concatenate([rng.integers(0, L[w], (n_ev[w], B)) for w in range(NW)]).astype(np.int32)
        cur_max = rp.prio.max(1).values.cpu().numpy()
        newp = rng.random((ev_tot, B)).astype(np.float32) * 2.0
        ev_w = np.repeat(np.arange(NW), n_ev)
        mode = step % 3
        for e in range(ev_tot):
            if mode == 0:   newp[e, ::7] = cur_max[ev_w[e]]                 # re-assert the maximum on some entries
            elif mode == 1: newp[e] *= 0.01                                  # everything far below: holders of the max disappear
            else:           newp[e, 3] = cur_max[ev_w[e]] + 1.0 + e          # a new, larger maximum
        ts, tp = torch.from_numpy(sidx).cuda(), torch.from_numpy(newp).cuda()
        _lib.check(vw.lib.rl_replay_update_prio(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs), B,
                                                C.c_void_p(ts.data_ptr()), C.c_void_p(tp.data_ptr()), vw._stream()))
        torch.cuda.synchronize()
        ref = before.clone()                                                  # sequential overwrite, later writes win
        ref = rp.prio.clone()                                                 # (values themselves are covered by test_learn_gpu)
        st = rp.maxst.cpu().numpy()
        for w in range(NW):
            if st[w, 1] > 0:
                known += 1
                m = float(ref[w].max())
                assert np.int32(st[w, 0]).view(np.float32) == np.float32(m) and st[w, 1] == int((ref[w] == m).sum()), (step, w, "update", mode)
        vw.update(); vw.top_up(40)
    assert known > 100


def _fake_events_scale(vw, rows, per_world):
    from reinlife_b200 import _lib as L
    cnt = torch.tensor(np.asarray(per_world), dtype=torch.int32)
    off = (torch.cumsum(cnt, 0) - cnt).int()
    k = L.ROWS_EVENT
    rows.count[k] = cnt.cuda(); rows.offset[k] = off.cuda(); rows.total[k] = int(cnt.sum())
    ids = [w * vw.S + e for w in range(vw.n_worlds) for e in range(int(per_world[w]))]
    if ids:
        rows.rows[k, :len(ids)] = torch.tensor(ids, dtype=torch.int32).cuda()
    return len(ids)


def _dqn_parts(d):
    return (("W1", 0, d.off_b1), ("b1", d.off_b1, d.off_w2t), ("W2", d.off_w2t, d.off_b2), ("b2", d.off_b2, d.off_wh),
            ("Wh", d.off_wh, d.off_bh), ("bh", d.off_bh, d.off_bh + 8))


@pytest.mark.parametrize("mode", ["dqn", "perdqn"])
def test_dqn_tensor_core_kernel_matches_fp32_kernel(mode):
    """k_learn_dqn_p<MODE> (tcgen05, fp16 operands, batch-major 128-row tiles, resident weight images and gradients) vs
    k_learn_dqn<MODE> (fp32 FMA, itself checked against the oracle above) at >= 5 tiles per persistent CTA, with skipped
    events, done rows and -- PERDQN -- per-event importance weights: summed gradients within 1 % of each tensor's gradient
    scale, per-event losses / errors within 2 %, the same event count."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, ReplayRings
    from reinlife_b200.Models import packing
    z, z1 = _golden2(), golden()
    rng = np.random.default_rng(23)
    B = 32 if mode == "dqn" else 64
    NW, cap, n_items = 64, 128, 100
    per_world = [48] * NW if mode == "dqn" else [24] * NW
    per_world[7] = 0
    vw, rows = _mk(NW)
    w0, tgt = _sd2(z, "train_dqn/w0"), _sd2(z, "train_dqn/target")
    rp = ReplayRings(NW, cap, "cuda", prioritized=False)
    _fill_rings(rp, rng, z1["obs"], n_items)
    n_ev = _fake_events(vw, rows, per_world)
    assert n_ev * B >= 5 * 128 * vw.lib.rl_learn_grid()
    sidx = rng.integers(0, n_items, size=(n_ev, B)).astype(np.int32)
    for e in rng.integers(0, n_ev, 25):
        sidx[e] = -1
    evw = torch.from_numpy(rng.uniform(0.2, 1.0, n_ev).astype(np.float32)).cuda()
    out = {}
    for kern in ("fp32", "tc"):
        brain = DeviceBrain(1, w0, "cuda", lr=5e-4, gamma=0.98, batch=B)
        brain.load_state_dict(tgt, target=True)
        brain.alloc_learn(rows.row_cap)
        brain.sample_idx[:n_ev] = torch.from_numpy(sidx).cuda()
        args = [C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs), C.c_void_p(brain.sample_idx.data_ptr())]
        if mode == "perdqn":
            args.append(C.c_void_p(evw.data_ptr()))
        args += [C.byref(brain.learn_bufs), vw._stream()]
        fn = {("dqn", "fp32"): vw.lib.rl_brain_learn_dqn, ("dqn", "tc"): vw.lib.rl_brain_learn_dqn_p,
              ("perdqn", "fp32"): vw.lib.rl_brain_learn_perdqn, ("perdqn", "tc"): vw.lib.rl_brain_learn_perdqn_p}[(mode, kern)]
        _lib.check(fn(*args))
        torch.cuda.synchronize()
        out[kern] = (brain.grad.cpu().numpy().copy(), brain.loss[:n_ev].cpu().numpy().copy(),
                     brain.new_prio.reshape(-1)[:n_ev * B].cpu().numpy().copy())
    d, m = packing.dims(1), packing.grad_mask(1)
    g32, l32, p32 = out["fp32"]
    gtc, ltc, ptc = out["tc"]
    nt = d.n_train
    assert g32[nt] == gtc[nt] == n_ev - len({int(e) for e in np.where(sidx[:, 0] < 0)[0]})
    for name, lo, hi in _dqn_parts(d):
        a, b = g32[lo:hi] * m[lo:hi], gtc[lo:hi] * m[lo:hi]
        scale = np.abs(a).max()
        assert scale > 0 and np.abs(a - b).max() / scale < 1e-2, (mode, name, np.abs(a - b).max() / scale, scale)
    valid = sidx[:, 0] >= 0
    np.testing.assert_allclose(ltc[valid], l32[valid], rtol=2e-2, atol=1e-3)
    if mode == "perdqn":
        v = np.repeat(valid, B)
        np.testing.assert_allclose(ptc[v], p32[v], rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("rule", ["dqn", "perdqn"])
def test_dqn_layout_act_tensor_core_matches_fp32_act(rule):
    """k_act_dqn_p (tcgen05 get_action of the DQN-layout brains, >= 3 tiles of 128 rows per persistent CTA) vs
    rl_brain_act_all (fp32 FMA, checked against the oracle in test_brain_gpu / test_perdqn_gpu): Q within 2e-2 of the Q scale,
    >= 99 % identical greedy actions, the same set of listed rows written, identical exploration draws and comparison rule."""
    from reinlife_b200 import _lib
    from reinlife_b200.Models import packing
    from reinlife_b200.World.vecworld import VecWorld
    from reinlife_b200.rows import RowLists
    z = _golden2()
    NW = 620
    vw = VecWorld(NW, 30, 30, 1, max_agents=100, seed=12)
    rows = RowLists(vw)
    vw.reset(); vw.top_up(100)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for _ in range(2):
        vw.set_actions(torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g))
        vw.step(); vw.update(); vw.top_up(100)
    rows.build(kinds_mask=1)
    n = int(rows.total[0])
    assert n >= 3 * 128 * vw.lib.rl_learn_grid()
    sd = _sd2(z, "train_dqn/w0")
    flat = torch.from_numpy(packing.pack(1, sd)).cuda()
    eps = torch.tensor([0.25], dtype=torch.float64, device="cuda")
    desc = (_lib.BrainAct * 1)(_lib.BrainAct(_lib.MODEL_DQN, _lib.ACT_DQN if rule == "dqn" else _lib.ACT_PERDQN, flat.data_ptr(), eps.data_ptr()))
    out = {}
    for kern in ("fp32", "tc"):
        q = torch.zeros((1, rows.row_cap, 8), device="cuda")
        vw.rec[:, :, 13] = 255
        if kern == "fp32":
            _lib.check(vw.lib.rl_brain_act_all(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), desc, 1, C.c_uint64(7),
                                               C.c_void_p(q.data_ptr()), None, vw._stream()))
        else:
            _lib.check(vw.lib.rl_brain_act_dqn_p(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), 0, C.byref(desc[0]), C.c_uint64(7),
                                                 C.c_void_p(q.data_ptr()), vw._stream()))
        torch.cuda.synchronize()
        out[kern] = (q[0, :n].cpu().numpy(), vw.rec[:, :, 13].cpu().numpy().view(np.int8).copy())
    q32, a32 = out["fp32"]
    qtc, atc = out["tc"]
    scale = np.abs(q32).max()
    assert scale > 0.1 and np.abs(q32 - qtc).max() < 2e-2 * scale
    assert (q32.argmax(1) == qtc.argmax(1)).mean() >= 0.99
    listed = a32 != -1
    assert listed.sum() == n and ((atc != -1) == listed).all()
    assert (a32[listed] == atc[listed]).mean() >= 0.99

"""GPU parity of replay + train events + Adam through the C ABI.

* one event == the reference's PERD3QNAgent.train(): weights after 3 consecutive Adam steps vs the reference
  (atol 1e-5), priorities written back (rtol 1e-4 / atol 1e-4);
* N events in one step == oracle 'mean of per-event gradients, one Adam step' (DESIGN.md learn-step semantics);
* the sampler == oracle.per_sample bit-for-bit (integer CDF); store == ring semantics of PERD3QN.py:143-155.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from brain_golden_util import golden, state_dict

pytestmark = pytest.mark.gpu


def _mk(n_worlds=1):
    from reinlife_b200.World.vecworld import VecWorld
    from reinlife_b200.rows import RowLists
    vw = VecWorld(n_worlds, 8, 8, 1, max_agents=100, seed=9)
    return vw, RowLists(vw)


def _fake_events(vw, rows, per_world):
    """Make the EVENT list by hand: world w has per_world[w] events (row ids w*S + e)."""
    import reinlife_b200._lib as L
    cnt = torch.tensor(per_world, dtype=torch.int32)
    off = (torch.cumsum(cnt, 0) - cnt).int()
    k = L.ROWS_EVENT
    rows.count[k] = cnt.cuda(); rows.offset[k] = off.cuda(); rows.total[k] = int(cnt.sum())
    ids = [w * vw.S + e for w in range(vw.n_worlds) for e in range(per_world[w])]
    rows.rows[k, :len(ids)] = torch.tensor(ids, dtype=torch.int32).cuda()
    return len(ids)


def _pad(x):
    return torch.from_numpy(np.pad(np.asarray(x, np.float32), ((0, 0), (0, 7))))


def _fill_ring(rp, w, obs, act, rew, nobs, done):
    n = len(act)
    rp.obs[w, :n] = _pad(obs).cuda(); rp.next_obs[w, :n] = _pad(nobs).cuda()
    rp.action[w, :n] = torch.from_numpy(np.asarray(act).astype(np.int8)).cuda()
    rp.reward[w, :n] = torch.from_numpy(np.asarray(rew).astype(np.float32)).cuda()
    rp.done[w, :n] = torch.from_numpy(np.asarray(done).astype(np.uint8)).cuda()
    rp.len[w] = n


def test_three_train_events_match_reference_train():
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, ReplayRings
    from reinlife_b200.Models import packing
    z = golden()
    vw, rows = _mk(1)
    brain = DeviceBrain(0, state_dict("train_perd3qn/w0"), "cuda", lr=1e-3, gamma=0.99)
    brain.load_state_dict(state_dict("train_perd3qn/target"), target=True)
    brain.alloc_learn(rows.row_cap)
    rp = ReplayRings(1, 512, "cuda")
    _fake_events(vw, rows, [1])
    for step in range(3):
        p = f"train_perd3qn/s{step}/"
        # ring position i holds the i-th sampled transition; the event samples positions 0..63 in order
        _fill_ring(rp, 0, z[p + "obs"], z[p + "action"], z[p + "reward"], z[p + "next_obs"], z[p + "done"])
        brain.sample_idx[0] = torch.arange(64, dtype=torch.int32).cuda()
        _lib.check(vw.lib.rl_brain_learn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                         C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs), vw._stream()))
        _lib.check(vw.lib.rl_brain_adam(C.byref(brain.learn_bufs), vw._stream()))
        torch.cuda.synchronize()
        got, want = brain.state_dict(), state_dict(p + "w")
        for k in want:
            np.testing.assert_allclose(got[k].numpy(), want[k], rtol=0, atol=1e-5, err_msg=f"step {step} {k}")
        ref_after, idx = z[p + "prio_after"], z[p + "indices"]
        last = {int(i): j for j, i in enumerate(idx)}
        newp = brain.new_prio[0].cpu().numpy()
        for i, j in last.items():
            np.testing.assert_allclose(newp[j], ref_after[i], rtol=1e-4, atol=1e-4)
    assert int(brain.adam_step) == 3
    flat = brain.params.cpu().numpy()
    m = packing.grad_mask(0)
    assert (flat[:len(m)][m == 0] == 0).all()                      # padding / structural zeros stay zero
    d = packing.dims(0)
    w2t = flat[d.off_w2t:d.off_b2].reshape(d.n1, d.n2)
    assert (flat[d.off_w2:d.off_w2 + d.n1 * d.n2].reshape(d.n2, d.n1) == w2t.T).all()   # output-major W2 copy


def test_batched_events_equal_mean_of_event_gradients():
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, ReplayRings
    from reinlife_b200.Models import packing
    from oracle import brain_oracle as bo
    z = golden()
    rng = np.random.default_rng(1)
    NW, per_world, cap = 5, [2, 0, 3, 1, 4], 256
    vw, rows = _mk(NW)
    w0, tgt = state_dict("train_perd3qn/w0"), state_dict("train_perd3qn/target")
    brain = DeviceBrain(0, w0, "cuda", lr=1e-3, gamma=0.99)
    brain.load_state_dict(tgt, target=True)
    brain.alloc_learn(rows.row_cap)
    rp = ReplayRings(NW, cap, "cuda")
    obs_all, rings = z["obs"], []
    for w in range(NW):
        n = 200
        o = obs_all[rng.integers(0, 512, n)]; no = obs_all[rng.integers(0, 512, n)]
        a = rng.integers(0, 8, n); r = rng.choice([0.0, 0.2, 0.5, -3.0, -20.0], n); d = (r < 0).astype(np.float64)
        _fill_ring(rp, w, o, a, r, no, d)
        rings.append((o, a, r, no, d))
    n_ev = _fake_events(vw, rows, per_world)
    sidx = rng.integers(0, 200, size=(n_ev, 64)).astype(np.int32)
    brain.sample_idx[:n_ev] = torch.from_numpy(sidx).cuda()
    _lib.check(vw.lib.rl_brain_learn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                     C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs), vw._stream()))
    torch.cuda.synchronize()
    events, e = [], 0
    for w in range(NW):
        o, a, r, no, d = rings[w]
        for _ in range(per_world[w]):
            i = sidx[e]; e += 1
            events.append((o[i], a[i], r[i], no[i], d[i]))
    g_ref, losses, prios = bo.dueling_batched_update(w0, tgt, events, 0.99)
    grad = brain.grad.cpu().numpy()
    nt = brain.dims.n_train
    assert grad[nt] == n_ev
    got = packing.unpack(0, np.concatenate([grad[:nt] / n_ev * packing.grad_mask(0), np.zeros(brain.dims.n_total - nt, np.float32)]))
    for k in g_ref:
        scale = max(1.0, np.abs(g_ref[k]).max())
        np.testing.assert_allclose(got[k].numpy(), g_ref[k], rtol=2e-4, atol=2e-5 * scale, err_msg=k)
    np.testing.assert_allclose(brain.loss[:n_ev].cpu().numpy(), np.array(losses), rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(brain.new_prio[:n_ev].cpu().numpy(), np.stack(prios), rtol=2e-4, atol=2e-4)


def test_store_sample_update_follow_the_reference_buffer():
    from reinlife_b200.brains import learn_step, DeviceBrain, ReplayRings
    from reinlife_b200.World.vecworld import VecWorld
    from reinlife_b200.rows import RowLists
    from reinlife_b200.Models import packing
    from oracle import brain_oracle as bo
    from oracle import ref_harness as rh
    NW, cap = 6, 96
    vw = VecWorld(NW, 12, 12, 2, max_agents=40, seed=21, world_id0=3)
    rows = RowLists(vw)
    vw.reset(); vw.top_up(40)
    torch.manual_seed(0)
    brains = [DeviceBrain(0, packing.default_init(0), "cuda", lr=1e-3, gamma=0.99) for _ in range(2)]
    for b in brains:
        b.alloc_learn(rows.row_cap)
    rps = [ReplayRings(NW, cap, "cuda") for _ in range(2)]
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    mirror = [[dict(prio=np.zeros(cap, np.float32), items=[None] * cap, len=0, pos=0) for _ in range(NW)] for _ in range(2)]
    tf = [3, 4]
    for step in range(8):
        obs_state = vw.obs_state.cpu().numpy().copy()
        vw.set_actions(torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g))
        vw.step()
        rows.build(kinds_mask=6, train_freq=tf, event_on=[1, 1])
        torch.cuda.synchronize()
        rec = vw.rec_host(); n = vw.n_agents.cpu().numpy()
        obs_prime = vw.obs_prime.cpu().numpy(); reward = vw.reward.cpu().numpy()
        for gene in range(2):
            learn_step(vw, rows, gene, brains[gene], rps[gene], vw.t)
            torch.cuda.synchronize()
            sidx = brains[gene].sample_idx.cpu().numpy(); newp = brains[gene].new_prio.cpu().numpy()
            ev_off = rows.offset[gene * 3 + 2].cpu().numpy()
            for w in range(NW):
                mr = mirror[gene][w]
                for s in range(n[w]):                                  # memorize (PERD3QN.py:143-155)
                    r = rec[w, s]
                    if r["gene"] != gene or r["age"] <= 1:
                        continue
                    maxp = mr["prio"].max() if mr["len"] > 0 else 1.0
                    mr["items"][mr["pos"]] = (obs_state[w, r["prev_slot"]], int(r["action"]), reward[w, s],
                                              obs_prime[w, s], bool(r["flags"] & 32))
                    mr["prio"][mr["pos"]] = maxp
                    mr["pos"] = (mr["pos"] + 1) % cap
                    mr["len"] = min(cap, mr["len"] + 1)
                assert int(rps[gene].len[w]) == mr["len"] and int(rps[gene].pos[w]) == mr["pos"]
                ev = [s for s in range(n[w]) if rec[w, s]["gene"] == gene and rec[w, s]["age"] > 1 and
                      (rec[w, s]["age"] % tf[gene] == 0 or rec[w, s]["flags"] & 32)]
                wts = bo.per_weight(mr["prio"][:mr["len"]])
                key = rh.world_key(21, 3 + w)
                for k_ev in range(len(ev)):                            # sampler: bit-exact integer CDF
                    u53 = [rh.draw(key, vw.t, rh.SITE["REPLAY_SAMPLE"], k_ev * 64 + i) >> 11 for i in range(64)]
                    assert sidx[ev_off[w] + k_ev].tolist() == bo.per_sample(wts, u53), (step, gene, w, k_ev)
                for k_ev in range(len(ev)):                            # update_priorities: sequential overwrite
                    for i in range(64):
                        mr["prio"][sidx[ev_off[w] + k_ev, i]] = newp[ev_off[w] + k_ev, i]
                assert (rps[gene].prio[w].cpu().numpy() == mr["prio"]).all(), (step, gene, w)
                assert (rps[gene].pw[w, :mr["len"]].cpu().numpy() == bo.per_weight(mr["prio"][:mr["len"]])).all()
                ro = rps[gene].obs[w].cpu().numpy(); rn = rps[gene].next_obs[w].cpu().numpy()
                ra = rps[gene].action[w].cpu().numpy(); rr = rps[gene].reward[w].cpu().numpy()
                rd = rps[gene].done[w].cpu().numpy()
                for p_ in range(mr["len"]):
                    it = mr["items"][p_]
                    assert (ro[p_] == it[0]).all() and ra[p_] == it[1] and rr[p_] == it[2]
                    assert (rn[p_] == it[3]).all() and rd[p_] == it[4]
        vw.update(); vw.top_up(40)
    assert all(int(b.adam_step) > 0 for b in brains)


def test_uniform_sampler_is_cpython_random_sample_on_the_ring():
    """rl_replay_sample_uniform == oracle.uniform_sample (== stdlib Random.sample, pinned on CPU) with population index 0 =
    oldest item of the deque (ring slot (pos - len + j) mod capacity), for both algorithm branches and a wrapped ring."""
    import reinlife_b200._lib as L
    from reinlife_b200.brains import ReplayRings
    from oracle import brain_oracle as bo
    from oracle import ref_harness as rh
    NW, cap = 6, 400
    vw, rows = _mk(NW)
    rp = ReplayRings(NW, cap, "cuda", prioritized=False)
    lens = [64, 100, 277, 278, 400, 30]
    poss = [64, 100, 277, 278, 123, 30]          # world 4: full ring that has wrapped (oldest item at slot 123)
    rp.len[:] = torch.tensor(lens, dtype=torch.int32).cuda(); rp.pos[:] = torch.tensor(poss, dtype=torch.int32).cuda()
    per_world = [2, 3, 1, 9, 11, 1]
    n_ev = _fake_events(vw, rows, per_world)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    for batch, it, n_iter, min_len in ((64, 0, 1, 0), (32, 3, 5, 90)):
        sidx = torch.full((rows.row_cap, batch), -7, dtype=torch.int32, device="cuda")
        status.zero_()
        L.check(vw.lib.rl_replay_sample_uniform(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs), batch, C.c_uint64(5),
                                                it, n_iter, min_len, C.c_void_p(sidx.data_ptr()), C.c_void_p(status.data_ptr()),
                                                vw._stream()))
        torch.cuda.synchronize()
        got = sidx.cpu().numpy()
        e = 0
        for w in range(NW):
            key = rh.world_key(9, w)
            for k_ev in range(per_world[w]):
                n = lens[w]
                if n <= min_len or n < batch:
                    assert (got[e] == -1).all()
                else:
                    j = bo.uniform_sample(n, batch, rh.uniform_sample_bits(key, 5, k_ev, it, n_iter))
                    first = (poss[w] - n) % cap
                    assert got[e].tolist() == [(first + x) % cap for x in j], (batch, w, k_ev)
                e += 1
        assert int(status) == (1 if batch == 64 else 0)      # world 5 holds 30 < 64 items: the reference raises ValueError


def test_three_d3qn_train_events_match_reference_train():
    """D3QNAgent.train() (Models/D3QN.py:97-116) x3 through the same event kernel (uniform replay, no priorities)."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, ReplayRings
    z = golden()
    vw, rows = _mk(1)
    brain = DeviceBrain(0, state_dict("train_d3qn/w0"), "cuda", lr=1e-3, gamma=0.99)
    brain.load_state_dict(state_dict("train_d3qn/target"), target=True)
    brain.alloc_learn(rows.row_cap)
    rp = ReplayRings(1, 512, "cuda", prioritized=False)
    _fake_events(vw, rows, [1])
    for step in range(3):
        p = f"train_d3qn/s{step}/"
        _fill_ring(rp, 0, z[p + "obs"], z[p + "action"], z[p + "reward"], z[p + "next_obs"], z[p + "done"])
        brain.sample_idx[0] = torch.arange(64, dtype=torch.int32).cuda()
        _lib.check(vw.lib.rl_brain_learn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                         C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs), vw._stream()))
        _lib.check(vw.lib.rl_brain_adam(C.byref(brain.learn_bufs), vw._stream()))
        torch.cuda.synchronize()
        got, want = brain.state_dict(), state_dict(p + "w")
        for k in want:
            np.testing.assert_allclose(got[k].numpy(), want[k], rtol=0, atol=1e-5, err_msg=f"step {step} {k}")


def _golden2():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "brain_golden2.npz"))


def _sd2(z, prefix):
    pre = prefix + "/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre) and "/" not in k[len(pre):]}


def test_dqn_five_iterations_match_reference_train():
    """train(q, q_target, memory, optimizer) (Models/DQN.py:142-153): the 5 (sample 32, smooth-L1, Adam) rounds of ONE
    train() call through rl_brain_learn_dqn + rl_brain_adam vs the weights the reference had after every optimizer step."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, ReplayRings
    z = _golden2()
    vw, rows = _mk(1)
    brain = DeviceBrain(1, _sd2(z, "train_dqn/w0"), "cuda", lr=5e-4, gamma=0.98, batch=32)
    brain.load_state_dict(_sd2(z, "train_dqn/target"), target=True)
    brain.alloc_learn(rows.row_cap)
    rp = ReplayRings(1, 64, "cuda", prioritized=False)
    _fake_events(vw, rows, [1])
    for it in range(5):
        p = f"train_dqn/i{it}/"
        _fill_ring(rp, 0, z[p + "obs"], z[p + "action"], z[p + "reward"], z[p + "next_obs"], 1.0 - z[p + "done_mask"])
        brain.sample_idx[0] = torch.arange(32, dtype=torch.int32).cuda()
        _lib.check(vw.lib.rl_brain_learn_dqn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                             C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs), vw._stream()))
        _lib.check(vw.lib.rl_brain_adam(C.byref(brain.learn_bufs), vw._stream()))
        torch.cuda.synchronize()
        got, want = brain.state_dict(), _sd2(z, p + "w")
        for k in want:
            np.testing.assert_allclose(got[k].numpy(), want[k], rtol=0, atol=1e-5, err_msg=f"iter {it} {k}")
    assert int(brain.adam_step) == 5


def test_dqn_batched_events_equal_mean_of_event_gradients_and_skip_short_rings():
    """N events in one iteration == oracle mean of per-event gradients; events flagged -1 by the sampler (ring <= 1000
    items, DQN.py:79) contribute nothing and are not counted."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, ReplayRings
    from reinlife_b200.Models import packing
    from oracle import brain_oracle as bo
    z, z1 = _golden2(), golden()
    rng = np.random.default_rng(3)
    NW, per_world, cap = 4, [2, 1, 0, 4], 128
    vw, rows = _mk(NW)
    w0, tgt = _sd2(z, "train_dqn/w0"), _sd2(z, "train_dqn/target")
    brain = DeviceBrain(1, w0, "cuda", lr=5e-4, gamma=0.98, batch=32)
    brain.load_state_dict(tgt, target=True)
    brain.alloc_learn(rows.row_cap)
    rp = ReplayRings(NW, cap, "cuda", prioritized=False)
    obs_all, rings = z1["obs"], []
    for w in range(NW):
        n = 100
        o = obs_all[rng.integers(0, 512, n)]; no = obs_all[rng.integers(0, 512, n)]
        a = rng.integers(0, 8, n); r = rng.choice([0.0, 0.2, 0.5, -3.0, -20.0], n); d = (r < 0).astype(np.float64)
        _fill_ring(rp, w, o, a, r, no, d)
        rings.append((o, a, r, no, d))
    n_ev = _fake_events(vw, rows, per_world)
    sidx = rng.integers(0, 100, size=(n_ev, 32)).astype(np.int32)
    skipped = {1, 5}
    for e in skipped:
        sidx[e] = -1
    brain.sample_idx[:n_ev] = torch.from_numpy(sidx).cuda()
    _lib.check(vw.lib.rl_brain_learn_dqn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                         C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs), vw._stream()))
    torch.cuda.synchronize()
    acc, losses, e = None, {}, 0
    for w in range(NW):
        o, a, r, no, d = rings[w]
        for _ in range(per_world[w]):
            if e not in skipped:
                i = sidx[e]
                g, loss = bo.dqn_iter_grads(w0, tgt, o[i], a[i], r[i], no[i], 1.0 - d[i])
                losses[e] = loss
                acc = g if acc is None else {k: acc[k] + g[k] for k in g}
            e += 1
    n_valid = n_ev - len(skipped)
    grad = brain.grad.cpu().numpy()
    nt = brain.dims.n_train
    assert grad[nt] == n_valid
    got = packing.unpack(1, np.concatenate([grad[:nt] / n_valid * packing.grad_mask(1), np.zeros(brain.dims.n_total - nt, np.float32)]))
    for k in acc:
        ref = acc[k] / n_valid
        np.testing.assert_allclose(got[k].numpy(), ref, rtol=2e-4, atol=2e-5 * max(1.0, np.abs(ref).max()), err_msg=k)
    loss_dev = brain.loss[:n_ev].cpu().numpy()
    for e, l in losses.items():
        np.testing.assert_allclose(loss_dev[e], l, rtol=2e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------ PPO
def _ppo_manual_plan(pd, rows, segs):
    """Lay segments (lists of transitions per world) into the data lists by hand and write the flat plan the way
    rl_ppo_store would.  segs[w] = list of segments, each a dict of arrays obs/next_obs/action/reward/prob_a/done."""
    import reinlife_b200._lib as L
    NW = len(segs)
    n_cons, flat, row_T, row_end, n_ev = [], [], [], [], 0
    cap = pd.traj.capacity
    for w, ss in enumerate(segs):
        j = 0
        for sg in ss:
            T = len(sg["action"])
            pd.traj.obs[w, j:j + T] = _pad(sg["obs"]).cuda(); pd.traj.next_obs[w, j:j + T] = _pad(sg["next_obs"]).cuda()
            pd.traj.action[w, j:j + T] = torch.from_numpy(np.asarray(sg["action"]).astype(np.int8)).cuda()
            pd.traj.reward[w, j:j + T] = torch.from_numpy(np.asarray(sg["reward"]).astype(np.float32)).cuda()
            pd.traj.prio[w, j:j + T] = torch.from_numpy(np.asarray(sg["prob_a"]).astype(np.float32)).cuda()
            pd.traj.done[w, j:j + T] = torch.from_numpy(np.asarray(sg["done"]).astype(np.uint8)).cuda()
            flat += [w * cap + j + k for k in range(T)]; row_T += [T] * T; row_end += [0] * (T - 1) + [1]
            j += T; n_ev += 1
        n_cons.append(j)
        pd.traj.len[w] = j
    off = np.concatenate([[0], np.cumsum(n_cons)]).astype(np.int32)
    pd.n_cons[:] = torch.tensor(n_cons, dtype=torch.int32).cuda(); pd.row_off[:] = torch.from_numpy(off).cuda()
    n = len(flat)
    pd.flat_src[:n] = torch.tensor(flat, dtype=torch.int32).cuda(); pd.row_T[:n] = torch.tensor(row_T, dtype=torch.int32).cuda()
    pd.row_end[:n] = torch.tensor(row_end, dtype=torch.uint8).cuda()
    rows.total[L.ROWS_EVENT] = n_ev
    return n_ev


@pytest.mark.parametrize("T", [1, 5, 37, 100])
def test_ppo_three_epochs_match_reference_learn(T):
    """PPO.learn() (Models/PPO.py:136-162) on one data list of T transitions: weights after each of the 3 optimizer
    steps vs the reference (golden minted by oracle/make_brain_golden2.py); T = 100 spans two 64-row tiles."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, PpoData
    z = _golden2()
    vw, rows = _mk(1)
    brain = DeviceBrain(2, _sd2(z, "train_ppo/w0"), "cuda", lr=5e-4, gamma=0.98, batch=64, has_target=False)
    brain.alloc_learn(rows.row_cap, need_batch_bufs=False)
    pd = PpoData(1, 128, 256, "cuda")
    p = f"train_ppo/T{T}/"
    seg = dict(obs=z[p + "obs"], next_obs=z[p + "next_obs"], action=z[p + "action"], reward=z[p + "reward"],
               prob_a=z[p + "prob_a"], done=z[p + "done"])
    _ppo_manual_plan(pd, rows, [[seg]])
    for e in range(3):
        _lib.check(vw.lib.rl_ppo_epoch(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(pd.bufs), C.byref(brain.learn_bufs), vw._stream()))
        _lib.check(vw.lib.rl_brain_adam(C.byref(brain.learn_bufs), vw._stream()))
        torch.cuda.synchronize()
        got, want = brain.state_dict(), _sd2(z, f"{p}e{e}")
        for k in want:   # Adam step 1 = lr*g/(|g|+1e-8): near-zero gradients amplify fp32 summation order
            np.testing.assert_allclose(got[k].numpy(), want[k], rtol=0, atol=3e-5, err_msg=f"T {T} epoch {e} {k}")
    assert int(brain.adam_step) == 3 and int(pd.status) == 0


def test_ppo_batched_segments_equal_mean_of_segment_gradients():
    """Several learn() calls of several worlds in one epoch == oracle mean over segments of the per-segment gradient
    (each a mean over its own T rows); td_target / delta / GAE per row vs the oracle."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, PpoData
    from reinlife_b200.Models import packing
    from oracle import brain_oracle as bo
    z, z1 = _golden2(), golden()
    rng = np.random.default_rng(5)
    w0 = _sd2(z, "train_ppo/w0")
    obs_all = z1["obs"]
    lens = [[3, 70], [], [1, 1, 20], [64]]
    vw, rows = _mk(len(lens))
    segs = []
    for ss in lens:
        cur = []
        for T in ss:
            o = obs_all[rng.integers(0, 512, T)]; no = obs_all[rng.integers(0, 512, T)]
            pi, _ = bo.ppo_forward(w0, o)
            a = rng.integers(0, 8, T)
            cur.append(dict(obs=o, next_obs=no, action=a, reward=rng.choice([0.0, 0.002, 0.005, -0.03, -0.42], T),
                            prob_a=(pi[np.arange(T), a] * rng.uniform(0.6, 1.4, T)).astype(np.float32), done=rng.random(T) < 0.2))
        segs.append(cur)
    brain = DeviceBrain(2, w0, "cuda", lr=5e-4, gamma=0.98, batch=64, has_target=False)
    brain.alloc_learn(rows.row_cap, need_batch_bufs=False)
    pd = PpoData(len(lens), 128, 512, "cuda")
    n_ev = _ppo_manual_plan(pd, rows, segs)
    _lib.check(vw.lib.rl_ppo_epoch(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(pd.bufs), C.byref(brain.learn_bufs), vw._stream()))
    torch.cuda.synchronize()
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
        np.testing.assert_allclose(got[k].numpy(), ref, rtol=5e-4, atol=2e-6 + 2e-5 * np.abs(ref).max(), err_msg=k)


def test_ppo_store_plan_and_compaction_follow_the_reference_data_list():
    """rl_ppo_store / rl_ppo_compact vs a python mirror of PPOAgent.learn's bookkeeping (PPO.py:71-77,113-134): the list
    order is the agent order, every trigger consumes the list so far, the tail stays for the next step."""
    import reinlife_b200._lib as L
    from reinlife_b200.brains import PpoData
    from reinlife_b200.World.vecworld import VecWorld
    from reinlife_b200.rows import RowLists
    NW, cap, tf = 5, 400, 4
    vw = VecWorld(NW, 10, 10, 2, max_agents=30, seed=4, world_id0=1)
    vw.enable_reward_div100()
    rows = RowLists(vw)
    vw.reset(); vw.top_up(30)
    pds = [PpoData(NW, cap, NW * cap, "cuda") for _ in range(2)]
    prob = torch.rand(NW * vw.S, device="cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    mirror = [[[] for _ in range(NW)] for _ in range(2)]
    for step in range(9):
        obs_state = vw.obs_state.cpu().numpy().copy()
        vw.set_actions(torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g))
        vw.step()
        rows.build(kinds_mask=6, train_freq=[tf, tf], event_on=[1, 1])
        torch.cuda.synchronize()
        rec = vw.rec_host(); n = vw.n_agents.cpu().numpy()
        obs_prime = vw.obs_prime.cpu().numpy(); r100 = vw.reward_div100.cpu().numpy(); rew = vw.reward.cpu().numpy()
        pr = prob.cpu().numpy().reshape(NW, vw.S)
        for gene in range(2):
            pd = pds[gene]
            L.check(vw.lib.rl_ppo_store(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), gene, C.c_void_p(prob.data_ptr()),
                                        tf, C.byref(pd.bufs), vw._stream()))
            torch.cuda.synchronize()
            exp_flat, exp_T, exp_end, n_ev = [], [], [], 0
            for w in range(NW):
                data = mirror[gene][w]
                cut = 0
                for s in range(n[w]):
                    r = rec[w, s]
                    if r["gene"] != gene or r["age"] <= 1:
                        continue
                    data.append((obs_state[w, r["prev_slot"]], int(r["action"]), r100[w, s], obs_prime[w, s],
                                 pr[w, r["prev_slot"]], bool(r["flags"] & 32)))
                    if r["age"] % tf == 0 or r["flags"] & 32:                      # learn(): consumes data[cut:len]
                        T = len(data) - cut
                        exp_flat += [w * cap + j for j in range(cut, len(data))]; exp_T += [T] * T; exp_end += [0] * (T - 1) + [1]
                        cut = len(data); n_ev += 1
                assert int(pd.traj.len[w]) == len(data) and int(pd.n_cons[w]) == cut
                to = pd.traj.obs[w, :len(data)].cpu().numpy(); tn = pd.traj.next_obs[w, :len(data)].cpu().numpy()
                ta = pd.traj.action[w].cpu().numpy(); tr = pd.traj.reward[w].cpu().numpy()
                tp = pd.traj.prio[w].cpu().numpy(); td = pd.traj.done[w].cpu().numpy()
                for j, it in enumerate(data):
                    assert (to[j] == it[0]).all() and ta[j] == it[1] and tr[j] == it[2] and (tn[j] == it[3]).all()
                    assert tp[j] == it[4] and td[j] == it[5]
                mirror[gene][w] = data[cut:]
            nr = int(pd.row_off[NW])
            assert nr == len(exp_flat) and n_ev == int(rows.total[gene * 3 + 2])
            assert pd.flat_src[:nr].cpu().tolist() == exp_flat and pd.row_T[:nr].cpu().tolist() == exp_T
            assert pd.row_end[:nr].cpu().tolist() == exp_end
            L.check(vw.lib.rl_ppo_compact(C.byref(vw.cfg), gene, C.byref(pd.bufs), vw._stream()))
            torch.cuda.synchronize()
            for w in range(NW):
                data = mirror[gene][w]
                assert int(pd.traj.len[w]) == len(data)
                to = pd.traj.obs[w, :len(data)].cpu().numpy(); tp = pd.traj.prio[w].cpu().numpy()
                for j, it in enumerate(data):
                    assert (to[j] == it[0]).all() and tp[j] == it[4]
            assert int(pd.status) == 0
        # reward / 100 is the float32 of the float64 quotient (PPO.py:73), not float32(reward) / 100
        nz = rew != 0
        assert (r100[nz] != 0).all()
        vw.update(); vw.top_up(30)
    assert sum(len(d) for g_ in mirror for d in g_) >= 0


def test_prioritized_sampler_tiles_match_oracle():
    """rl_replay_sample at the default capacity: the coalesced register-tile scan (several tiles of 1024 weights, ragged lengths,
    a length of 1, a full ring) and the chunked fallback (capacity not a multiple of 4) draw exactly the oracle's indices."""
    from reinlife_b200 import _lib
    from reinlife_b200.brains import ReplayRings
    from oracle import brain_oracle as bo
    from oracle import ref_harness as rh
    for cap, lens in ((10000, [10000, 1, 1023, 1024, 1025, 4097, 9999, 3]), (1003, [1003, 7, 512, 1000])):
        NW = len(lens)
        vw, rows = _mk(NW)
        rp = ReplayRings(NW, cap, "cuda")
        rng = np.random.default_rng(cap)
        prio = (rng.random((NW, cap)) ** 4 * 3.0).astype(np.float32)
        prio[0, ::7] = 0.0                                     # zero-weight entries are never drawn
        pw = np.stack([bo.per_weight(p) for p in prio])
        rp.prio.copy_(torch.from_numpy(prio)); rp.pw.copy_(torch.from_numpy(pw))
        rp.len.copy_(torch.tensor(lens, dtype=torch.int32))
        per_world = [2, 1, 1, 1, 2, 1, 1, 1][:NW]
        n_ev = _fake_events(vw, rows, per_world)
        sidx = torch.full((rows.row_cap, 64), -7, dtype=torch.int32, device="cuda")
        t = 5
        _lib.check(vw.lib.rl_replay_sample(C.byref(vw.cfg), C.byref(rows.bufs), C.c_int32(0), C.byref(rp.bufs), C.c_int32(64),
                                           C.c_uint64(t), C.c_void_p(sidx.data_ptr()), vw._stream()))
        torch.cuda.synchronize()
        got = sidx.cpu().numpy()
        ev = 0
        for w in range(NW):
            key = rh.world_key(9, w)
            for k_ev in range(per_world[w]):
                u53 = [rh.draw(key, t, rh.SITE["REPLAY_SAMPLE"], k_ev * 64 + i) >> 11 for i in range(64)]
                assert got[ev].tolist() == bo.per_sample(pw[w, :lens[w]], u53), (cap, w, k_ev)
                ev += 1
        assert ev == n_ev

"""GPU parity of the PERDQN brain (Models/PERDQN.py) through the C ABI.

* the recorded run of the reference's PERDQNAgent (tests/golden/brain_golden3.npz: 150 stores, 3 train_model calls, 250
  stores across the ring wrap, 1 train_model) replayed on the device: SumTree bit-exact (float64 patterns) after every
  phase, weights after every optimizer step atol 1e-5, errors rtol 1e-4;
* the stratified sampler == the pinned oracle on the counter RNG: slots exact, importance-weight mean rtol 1e-6, beta
  exact, short memories skipped;
* get_action: forward vs the pinned oracle (rtol 1e-4 / atol 1e-4), `u <= eps` rule exact;
* trainer() with PERDQN brains: learns, epsilon steps once per optimizer step, target == model after the last trigger.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from perdqn_golden_util import golden3, meta3, sd3, transitions

pytestmark = pytest.mark.gpu


def _mk(n_worlds=1, seed=9):
    from reinlife_b200.World.vecworld import VecWorld
    from reinlife_b200.rows import RowLists
    vw = VecWorld(n_worlds, 8, 8, 1, max_agents=100, seed=seed)
    return vw, RowLists(vw)


def _fake_rows(vw, rows, kind, per_world):
    cnt = torch.tensor(per_world, dtype=torch.int32)
    off = (torch.cumsum(cnt, 0) - cnt).int()
    rows.count[kind] = cnt.cuda(); rows.offset[kind] = off.cuda(); rows.total[kind] = int(cnt.sum())
    ids = [w * vw.S + e for w in range(vw.n_worlds) for e in range(per_world[w])]
    rows.rows[kind, :len(ids)] = torch.tensor(ids, dtype=torch.int32).cuda()
    return len(ids)


def _pad(x):
    return torch.from_numpy(np.pad(np.asarray(x, np.float32), ((0, 0), (0, 7))))


def test_recorded_reference_run_replays_on_the_device():
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, ReplayRings, SumTrees
    from reinlife_b200.Models import packing
    z, M = golden3(), meta3()
    cap = M["capacity"]
    vw, rows = _mk(1)
    lib, st = vw.lib, vw._stream()
    brain = DeviceBrain(packing.PERDQN, sd3("run/w0"), "cuda", lr=M["lr"], gamma=M["gamma"], batch=64)
    brain.load_state_dict(sd3("run/target"), target=True)
    brain.alloc_learn(rows.row_cap)
    rp = ReplayRings(1, cap, "cuda", prioritized=False)
    mem = SumTrees(1, cap, "cuda", train_start=M["train_start"], ev_cap=rows.row_cap)
    assert np.float64(np.float32(mem.p_new)) == z["run/add_leaf"][0]            # float32(0.01) ** 0.6 as torch computes it
    eps = torch.ones(1, dtype=torch.float64, device="cuda")
    state = {"n": 0, "upd": 0, "train": 0}
    s_all, a_all, r_all, s2_all, d_all = transitions(0, len(z["run/action"]))

    def store(n):                       # Memory.add x n through rl_sumtree_add, <= 40 STORE rows per call
        while n:
            k = min(n, 40)
            _fake_rows(vw, rows, _lib.ROWS_STORE, [k])
            _lib.check(lib.rl_sumtree_add(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs), C.byref(mem.bufs), st))
            i = np.arange(state["n"], state["n"] + k)           # the ring itself (rl_replay_store is covered in test_learn_gpu)
            slots = torch.from_numpy(i % cap).cuda()
            rp.obs[0, slots] = _pad(s_all[i]).cuda(); rp.next_obs[0, slots] = _pad(s2_all[i]).cuda()
            rp.action[0, slots] = torch.from_numpy(a_all[i].astype(np.int8)).cuda()
            rp.reward[0, slots] = torch.from_numpy(r_all[i].astype(np.float32)).cuda()
            rp.done[0, slots] = torch.from_numpy(d_all[i].astype(np.uint8)).cuda()
            state["n"] += k
            rp.pos[0] = state["n"] % cap; rp.len[0] = min(cap, state["n"])
            n -= k

    def check_tree(name):
        torch.cuda.synchronize()
        got, want = mem.tree[0].cpu().numpy(), z[f"run/{name}/tree"]
        assert np.array_equal(got, want), (name, int((got != want).sum()), np.abs(got - want).max())

    def train():
        k = state["train"]
        idxs, isw = z[f"run/sample{k}/idx"], z[f"run/sample{k}/isw"]
        _fake_rows(vw, rows, _lib.ROWS_EVENT, [1])
        brain.sample_idx[0] = torch.from_numpy((idxs - cap + 1).astype(np.int32)).cuda()
        mem.ev_weight[0] = float(isw.astype(np.float32).sum(dtype=np.float32) / np.float32(64))
        _lib.check(lib.rl_brain_learn_perdqn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                             C.c_void_p(brain.sample_idx.data_ptr()), C.c_void_p(mem.ev_weight.data_ptr()),
                                             C.byref(brain.learn_bufs), st))
        _lib.check(lib.rl_brain_adam(C.byref(brain.learn_bufs), st))
        _lib.check(lib.rl_perdqn_epsilon_step(C.byref(brain.learn_bufs), C.c_void_p(eps.data_ptr()),
                                              C.c_double(M["eps_min"]), C.c_double(M["eps_decay"]), st))
        torch.cuda.synchronize()
        ref_err = z["run/upd_err"][state["upd"]:state["upd"] + 64]
        np.testing.assert_allclose(brain.new_prio[0].cpu().numpy(), ref_err, rtol=1e-4, atol=2e-5)
        got, want = brain.state_dict(), sd3(f"run/step{k}")
        for name in want:
            np.testing.assert_allclose(got[name].numpy(), want[name], rtol=0, atol=1e-5, err_msg=f"train {k} {name}")
        # Memory.update with the reference's own float32 errors: the tree must come out bit-identical
        brain.new_prio[0] = torch.from_numpy(ref_err).cuda()
        _lib.check(lib.rl_sumtree_update(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs), C.byref(mem.bufs),
                                         C.c_int32(64), C.c_void_p(brain.sample_idx.data_ptr()),
                                         C.c_void_p(brain.new_prio.data_ptr()), st))
        state["upd"] += 64; state["train"] += 1

    store(150); check_tree("p0_store150")
    for k in range(3):
        train(); check_tree(f"p{k + 1}_train")
    store(250); check_tree("p4_store250")
    train(); check_tree("p5_train")
    assert int(brain.adam_step) == 4
    assert abs(float(eps) - z["run/p5_train/scal"][3]) < 1e-12
    # the padded hidden units never became parameters
    d = brain.dims
    flat = brain.params.cpu().numpy()
    assert not flat[0:d.off_b1].reshape(160, d.n1)[:, 64:].any() and not flat[d.off_w2t:d.off_b2].reshape(d.n1, d.n2)[64:].any()


def test_sampler_matches_oracle_on_the_counter_rng():
    from reinlife_b200 import _lib
    from reinlife_b200.brains import ReplayRings, SumTrees
    from oracle import brain_oracle as bo
    from oracle import ref_harness as rh
    rng = np.random.default_rng(4)
    NW, cap, t = 5, 37, 11
    vw, rows = _mk(NW, seed=21)
    rp = ReplayRings(NW, cap, "cuda", prioritized=False)
    mem = SumTrees(NW, cap, "cuda", train_start=10, ev_cap=rows.row_cap)
    fills = [37, 20, 5, 12, 37]                 # world 2 holds fewer than train_start items: its events are skipped
    per_world = [1, 3, 2, 0, 2]
    trees = []
    for w in range(NW):
        tr = bo.SumTreeOracle(cap)
        for _ in range(fills[w] + (9 if fills[w] == cap else 0)):      # full memories have wrapped
            tr.add(bo.perdqn_priority(np.float32(0.0)))
        for _ in range(3 * fills[w]):                                   # spread the priorities over two decades
            tr.update(int(rng.integers(fills[w])) + cap - 1, bo.perdqn_priority(np.float32(rng.choice([0.0, 0.3, 2.5, 40.0]) * rng.random())))
        trees.append(tr)
        mem.tree[w] = torch.from_numpy(tr.tree).cuda()
        rp.len[w] = tr.n_entries; rp.pos[w] = tr.write
    n_ev = _fake_rows(vw, rows, _lib.ROWS_EVENT, per_world)
    sidx = torch.full((rows.row_cap, 64), -7, dtype=torch.int32, device="cuda")
    _lib.check(vw.lib.rl_sumtree_sample(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs), C.byref(mem.bufs),
                                        C.c_int32(64), C.c_uint64(t), C.c_void_p(sidx.data_ptr()),
                                        C.c_void_p(mem.ev_weight.data_ptr()), vw._stream()))
    torch.cuda.synchronize()
    assert int(mem.status) == 0
    got, gw, ev = sidx.cpu().numpy(), mem.ev_weight.cpu().numpy(), 0
    for w in range(NW):
        key = rh.world_key(21, w)
        for e in range(per_world[w]):
            if fills[w] < 10:
                assert (got[ev] == -1).all() and gw[ev] == 0.0
            else:
                u = lambda i, tries, e=e: rh.uniform(rh.draw(key, t, rh.SITE["SUMTREE_SAMPLE"], (e * 64 + i) * 64 + tries))
                slots, _, isw = trees[w].sample(64, indexed_u=u)
                assert got[ev].tolist() == slots, (w, e)
                np.testing.assert_allclose(gw[ev], isw.astype(np.float32).mean(dtype=np.float64), rtol=1e-6)
            ev += 1
        assert float(mem.beta[w]) == trees[w].beta                       # +0.001 per sample() call of that world
    assert ev == n_ev and (got[n_ev:] == -7).all()


def test_get_action_matches_oracle_forward_and_rule():
    from reinlife_b200 import _lib
    from reinlife_b200.Models import packing
    from oracle import brain_oracle as bo
    from oracle import ref_harness as rh
    from brain_golden_util import golden
    from test_brain_gpu import _inject_obs, _setup_world
    sd = sd3("fwd/w")
    obs = golden()["obs"]
    vw, rows = _setup_world()
    per_row = _inject_obs(vw, obs)
    rows.build(kinds_mask=1)
    flat = torch.from_numpy(packing.pack(packing.PERDQN, sd)).cuda()
    G, eps, t_act = vw.G, 0.3, 7
    eps_dev = torch.full((G,), eps, dtype=torch.float64, device="cuda")
    acts = (_lib.BrainAct * G)(*[_lib.BrainAct(_lib.MODEL_DQN, _lib.ACT_PERDQN, flat.data_ptr(), eps_dev.data_ptr() + 8 * g)
                                 for g in range(G)])
    q_out = torch.zeros((G, rows.row_cap, 8), device="cuda")
    _lib.check(vw.lib.rl_brain_act_all(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), acts, G, C.c_uint64(t_act),
                                       C.c_void_p(q_out.data_ptr()), None, vw._stream()))
    torch.cuda.synchronize()
    rec, checked, explored = vw.rec_host(), 0, 0
    q_ref = bo.perdqn_forward(sd, obs)
    for g in range(G):
        lst = rows.list(g, 0)
        q = q_out[g, :len(lst)].cpu().numpy()
        np.testing.assert_allclose(q, np.stack([q_ref[per_row[int(r)]] for r in lst], 0), rtol=1e-4, atol=1e-4)
        for i, r in enumerate(lst):
            w, s = divmod(int(r), vw.S)
            key = rh.world_key(3, 40 + w)
            u = rh.uniform(rh.draw(key, t_act, rh.SITE["ACT_EXPLORE"], s))
            rb = rh.below(rh.draw(key, t_act, rh.SITE["ACT_RANDOM"], s), 8)
            assert rec[w, s]["action"] == bo.perdqn_rule(q[i], eps, u, rb), (g, i)
            checked += 1; explored += u <= eps
    assert checked == int(vw.n_agents.sum()) and 0 < explored < checked
    # the reference's per-agent plugin call on a host observation (Helpers/tester.py:66-68): same network, epsilon 0
    from reinlife_b200.Models import PERDQN
    b = PERDQN(training=False)
    b.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    for i in (0, 5, 17):
        assert b.get_action(obs[i]) == int(np.argmax(q_ref[i]))


def test_trainer_learns_perdqn():
    """PERDQN trains on the device: one optimizer step per step with a trained trigger, epsilon one decay per optimizer
    step (PERDQN.py:132-133), target_model == model after the last trigger (PERDQN.py:195), memories wrap."""
    import reinlife_b200 as rl
    from reinlife_b200.Models import PERDQN
    torch.manual_seed(4)
    brains = [PERDQN(train_freq=5, capacity=300, explore_step=100), PERDQN(train_freq=5, capacity=300, explore_step=100)]
    for b in brains:
        b.train_start = 120
    w0 = [b.model.state_dict()["fc.0.weight"].clone() for b in brains]
    env = rl.trainer(brains, n_episodes=40, width=12, height=12, max_agents=30, update_interval=10, print_results=False,
                     save=False, n_worlds=6, seed=8, saturate_to=30)
    torch.cuda.synchronize()
    for g, b in enumerate(brains):
        steps = int(b._dev.adam_step)
        assert steps > 0 and torch.isfinite(b._dev.params).all()
        assert not torch.equal(b.model.state_dict()["fc.0.weight"], w0[g])
        assert b.model.state_dict()["fc.0.weight"].shape == (64, 153)
        for k, v in b.model.state_dict().items():
            assert torch.equal(v, b.target_model.state_dict()[k])
        assert abs(env.epsilons()[g] - (1.0 - steps * b.epsilon_decay)) < 1e-9
        assert int(b._replay.len.max()) == 300 and int(b.memory.status) == 0
        tree = b.memory.tree.cpu().numpy()
        leaves = tree[:, 299:]
        assert np.allclose(tree[:, 0], leaves.sum(1), rtol=1e-5) and (leaves[b._replay.len.cpu().numpy() == 300] > 0).all()
        assert float(b.memory.beta.max()) > 0.4

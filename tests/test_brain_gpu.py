"""GPU parity of the brain kernels through the C ABI.

Tolerances (north_star: 'stated fp tolerance for network outputs'): network outputs rtol 1e-4 / atol 1e-4
against the reference modules' float32 outputs (|Q| is O(10-50) with the pretrained weights; fp32 FMA order
differs from MKL's); action selection is integer work and must be exact given the kernel's own outputs.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from brain_golden_util import golden, state_dict

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-4, atol=1e-4)


def _setup_world(n_worlds=8, target=70, G=2, seed=3):
    from reinlife_b200.World.vecworld import VecWorld
    from reinlife_b200.rows import RowLists
    vw = VecWorld(n_worlds, 30, 30, G, max_agents=100, seed=seed, world_id0=40)
    vw.reset(); vw.top_up(target)
    return vw, RowLists(vw)


def _inject_obs(vw, obs):
    """Overwrite the observation rows of the listed agents with golden observations (row order = list order)."""
    n = vw.n_agents.cpu().numpy()
    k = 0
    host = torch.zeros_like(vw.obs_state, device="cpu")
    per_row = {}
    for w in range(vw.n_worlds):
        for s in range(n[w]):
            host[w, s, :153] = torch.from_numpy(obs[k % len(obs)].astype(np.float32))
            per_row[w * vw.S + s] = k % len(obs)
            k += 1
    vw.obs_state.copy_(host)
    return per_row


@pytest.mark.parametrize("name,kind,rule,eps", [("perd3qn", 0, 0, 0.3), ("d3qn", 0, 0, 0.0), ("dqn", 1, 1, 0.2), ("ppo", 2, 2, 0.0)])
def test_act_matches_reference_forward_and_rule(name, kind, rule, eps):
    from reinlife_b200 import _lib
    from reinlife_b200.Models import packing
    from oracle import ref_harness as rh   # pure-python RNG spec only (no reference import)
    z = golden()
    vw, rows = _setup_world()
    per_row = _inject_obs(vw, z["obs"])
    rows.build(kinds_mask=1)
    flat = torch.from_numpy(packing.pack(kind, state_dict(name))).cuda()
    G = vw.G
    eps_dev = torch.full((G,), eps, dtype=torch.float64, device="cuda")
    acts = (_lib.BrainAct * G)(*[_lib.BrainAct(kind, rule, flat.data_ptr(), eps_dev.data_ptr() + 8 * g) for g in range(G)])
    q_out = torch.zeros((G, rows.row_cap, 8), device="cuda")
    prob_out = torch.zeros(vw.n_worlds * vw.S, device="cuda")
    t_act = 5
    _lib.check(vw.lib.rl_brain_act_all(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), acts, G,
                                       C.c_uint64(t_act), C.c_void_p(q_out.data_ptr()), C.c_void_p(prob_out.data_ptr()),
                                       vw._stream()))
    torch.cuda.synchronize()
    rec = vw.rec_host()
    ref_key = {"perd3qn": "perd3qn_q_rows", "d3qn": "d3qn_q_rows", "dqn": "dqn_q_rows", "ppo": "ppo_pi_rows"}[name]
    checked = 0
    for g in range(G):
        lst = rows.list(g, 0)
        q = q_out[g, :len(lst)].cpu().numpy()
        want = np.stack([z[ref_key][per_row[int(r)]] for r in lst], 0)
        np.testing.assert_allclose(q, want, **(dict(rtol=1e-4, atol=1e-6) if name == "ppo" else TOL))
        for i, r in enumerate(lst):
            w, s = divmod(int(r), vw.S)
            assert rec[w, s]["gene"] == g
            key = rh.world_key(3, 40 + w)
            u = rh.uniform(rh.draw(key, t_act, rh.SITE["ACT_EXPLORE"], s))
            rb = rh.below(rh.draw(key, t_act, rh.SITE["ACT_RANDOM"], s), 8)
            if rule == 0:
                a = int(np.argmax(q[i])) if u > eps else rb
            elif rule == 1:
                a = rb if u < eps else int(np.argmax(q[i]))
            else:
                us = rh.uniform(rh.draw(key, t_act, rh.SITE["ACT_SAMPLE"], s))
                c, a = np.float32(0), 7
                for j in range(8):
                    c = np.float32(c + q[i][j])
                    if us < float(c):
                        a = j
                        break
                assert prob_out[int(r)].item() == q[i][a]
            assert rec[w, s]["action"] == a, (name, g, i)
            checked += 1
    assert checked == int(vw.n_agents.sum())


def test_row_lists_are_world_major_and_filtered():
    vw, rows = _setup_world(n_worlds=5, target=60, G=3)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for _ in range(3):
        vw.set_actions(torch.randint(0, 8, (5, vw.S), device="cuda", dtype=torch.int8, generator=g))
        vw.step()
    tf = [20, 5, 3]
    rows.build(kinds_mask=7, train_freq=tf, event_on=[1, 1, 0])
    torch.cuda.synchronize()
    rec = vw.rec_host(); n = vw.n_agents.cpu().numpy()
    for gene in range(3):
        want = {0: [], 1: [], 2: []}
        for w in range(5):
            for s in range(n[w]):
                r = rec[w, s]
                if r["gene"] != gene:
                    continue
                want[0].append(w * vw.S + s)
                if r["age"] > 1:
                    want[1].append(w * vw.S + s)
                    if [1, 1, 0][gene] and (r["age"] % tf[gene] == 0 or r["flags"] & 32):
                        want[2].append(w * vw.S + s)
        for kind in range(3):
            assert rows.list(gene, kind).tolist() == want[kind], (gene, kind)

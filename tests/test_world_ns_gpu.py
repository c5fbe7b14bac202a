"""Non-static families on the device, World side (SURVEY 8 rows a17 / a18 non-static branches, a21 _update_best_agents):
k_world_reset / step / update <NS> vs whole trajectories of the UNMODIFIED reference run with static_families=False
(tests/golden/world_ns_golden.npz, minted by oracle/make_golden_ns.py) -- after every step() and update_env(): cell types,
agent list, float32(reward), Agent.fitness (float64 bit patterns), object identity, the ten best agents (identity,
fitness, brain id), max_gene and the _produce event -- and vs the C oracle on a batch of larger worlds."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from golden_util import rec_diff, rec_equal   # noqa: E402

pytestmark = pytest.mark.gpu
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "world_ns_golden.npz")
FIELDS = ("cell", "health", "age", "max_age", "gene", "flags", "prev_slot")


def _best(st):
    return [b.serial for b in st.best], [b.fitness for b in st.best], [b.brain for b in st.best]


@pytest.mark.parametrize("ti", range(4))
def test_non_static_trajectory_matches_reference(ti):
    from reinlife_b200.World.vecworld import VecWorld
    z = np.load(PATH)
    m = json.loads(bytes(z["meta"]).decode())[ti]
    g = lambda k: z[f"t{ti}_{k}"]   # noqa: E731
    vw = VecWorld(1, m["height"], m["width"], m["n_genes"], max_agents=m["max_agents"], seed=m["seed"], world_id0=m["world"],
                  static_families=False)
    vw.reset()
    torch.cuda.synchronize()
    n = int(vw.n_agents[0])
    assert (vw.type[0].cpu().numpy() == g("reset_type").reshape(-1)).all() and rec_equal(vw.rec_host()[0, :n], g("reset_rec"), FIELDS)
    st = vw.ns_host()[0]
    assert st.max_gene == m["n_genes"] and _best(st)[0] == [-1 - k for k in range(10)]
    assert vw.serial[0, :n].cpu().tolist() == list(range(n)) and float(vw.fitness[0, :n].abs().sum()) == 0.0
    off = dict(a=0, s=0, u=0)
    for t in range(m["steps"]):
        acts = g("actions")[off["a"]:off["a"] + n]
        a = np.zeros((1, vw.S), np.int8); a[0, :n] = acts
        vw.set_actions(a)
        off["a"] += n
        for phase, key in (("step", "s"), ("upd", "u")):
            vw.step() if phase == "step" else vw.update()
            torch.cuda.synchronize()
            n = int(g(phase + "_n")[t])
            assert int(vw.n_agents[0]) == n, (t, phase)
            lo = off[key]
            assert (vw.type[0].cpu().numpy() == g(phase + "_type")[t].reshape(-1)).all(), (t, phase)
            want = g(phase + "_rec")[lo:lo + n]
            got = vw.rec_host()[0, :n]
            assert rec_equal(got, want, FIELDS), (t, phase, rec_diff(got, want, FIELDS))
            assert (vw.fitness[0, :n].cpu().numpy().view(np.uint64) == g(phase + "_fitness")[lo:lo + n].view(np.uint64)).all(), (t, phase)
            assert (vw.serial[0, :n].cpu().numpy() == g(phase + "_serial")[lo:lo + n]).all(), (t, phase)
            st = vw.ns_host()[0]
            bs, bf, bb = _best(st)
            assert bs == g(phase + "_best_serial")[t].tolist(), (t, phase)
            assert np.array_equal(np.array(bf).view(np.uint64), g(phase + "_best_fitness")[t].view(np.uint64)), (t, phase)
            assert bb == g(phase + "_best_brain")[t].tolist(), (t, phase)
            if phase == "step":
                assert (vw.reward[0, :n].cpu().numpy() == g("step_reward")[lo:lo + n].astype(np.float32)).all(), t
            else:
                assert st.max_gene == int(g("upd_max_gene")[t])
                pg, pk = g("upd_produced")[t]
                assert (st.produced_gene, st.produced_src_best) == (int(pg), int(pk)), t
                if pg >= 0:
                    assert st.produced_src_brain == int(g("upd_best_brain")[t][pk]), t
                assert int(vw.n_lineages[0]) == len(set(got["gene"].tolist())), t
            off[key] += n


def test_non_static_batch_matches_oracle_60x60():
    """48 worlds of 60x60 (BASELINE configs[4] grid) for 150 steps with random actions: every world == the C oracle's
    sequential restatement (types, records, rewards, observations, fitness, serials, best tables, max_gene, events)."""
    from reinlife_b200.World.vecworld import VecWorld
    from oracle.world_oracle import OracleWorlds
    NW, H, W, G = 48, 60, 60, 2
    vw = VecWorld(NW, H, W, G, max_agents=400, seed=17, world_id0=5, static_families=False, slot_cap=1024)
    ow = OracleWorlds(NW, H, W, G, max_agents=400, seed=17, world_id0=5, static_families=False, slot_cap=1024)
    vw.reset(); ow.reset()
    gen = torch.Generator(device="cuda"); gen.manual_seed(4)
    for t in range(150):
        acts = torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=gen)
        vw.set_actions(acts); ow.set_actions(acts.cpu().numpy())
        vw.step(); ow.step()
        if t % 10 == 9 or t < 3:
            _compare(vw, ow, NW, step=True)
        vw.update(); ow.update()
        if t % 10 == 9 or t < 3:
            _compare(vw, ow, NW, step=False)
    _compare(vw, ow, NW, step=False)
    assert int(vw.status.max()) == 0
    assert max(s.max_gene for s in vw.ns_host()) > G + 3


def _compare(vw, ow, NW, step):
    torch.cuda.synchronize()
    rec, typ, n = vw.rec_host(), vw.type.cpu().numpy(), vw.n_agents.cpu().numpy()
    fit, ser = vw.fitness.cpu().numpy(), vw.serial.cpu().numpy()
    obs = (vw.obs_prime if step else vw.obs_state).cpu().numpy()
    rew = vw.reward.cpu().numpy()
    st = vw.ns_host()
    for w in range(NW):
        k = int(ow.n[w])
        assert n[w] == k and (typ[w] == ow.type[w]).all(), w
        assert rec_equal(rec[w, :k], ow.rec[w, :k], FIELDS), (w, rec_diff(rec[w, :k], ow.rec[w, :k], FIELDS))
        assert (fit[w, :k].view(np.uint64) == ow.fitness[w, :k].view(np.uint64)).all() and (ser[w, :k] == ow.serial[w, :k]).all(), w
        assert (obs[w, :k, :153] == ow.obs[w, :k].astype(np.float32)).all(), w
        if step:
            assert (rew[w, :k] == ow.reward[w, :k].astype(np.float32)).all(), w
        o = ow.ns[w]
        assert (st[w].max_gene, st[w].next_serial) == (o.max_gene, o.next_serial), w
        assert (st[w].produced_gene, st[w].produced_src_best, st[w].produced_src_brain) == (o.produced_gene, o.produced_src_best, o.produced_src_brain), w
        for a, b in zip(st[w].best, o.best):
            assert (a.serial, a.brain) == (b.serial, b.brain) and np.float64(a.fitness).view(np.uint64) == np.float64(b.fitness).view(np.uint64), w

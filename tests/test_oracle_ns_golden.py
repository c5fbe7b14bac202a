"""Pins the non-static-families part of the C oracle (oracle/rl_oracle.c: rlo_*_ns) against whole trajectories of the
unmodified reference run with static_families=False (tests/golden/world_ns_golden.npz, oracle/make_golden_ns.py):
after every step() and every update_env() the cell types, agent list, float64 rewards, Agent.fitness, object identity,
the ten best agents (identity, fitness, brain id), max_gene and the _produce event must be identical.

This is the World side of SURVEY 8 rows a17/a18 (non-static branches) and a21 (_update_best_agents); the CUDA path does
not implement non-static families yet -- the oracle is the pinned specification for it."""
import json
import os

import numpy as np
import pytest

from golden_util import rec_diff, rec_equal
from oracle.world_oracle import OracleWorlds

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "world_ns_golden.npz")
FIELDS = ("cell", "health", "age", "max_age", "gene", "flags", "prev_slot")


def _load():
    z = np.load(PATH)
    return z, json.loads(bytes(z["meta"]).decode())


def _best(o):
    b = o.ns[0].best
    return [x.serial for x in b], [x.fitness for x in b], [x.brain for x in b]


def test_golden_trajectories_exercise_the_non_static_rules():
    _, meta = _load()
    assert len(meta) == 4 and all(m["produced"] >= 5 and m["best_replaced"] >= 5 for m in meta)
    assert all(m["final_max_gene"] == m["n_genes"] + m["produced"] for m in meta)      # gene = ++max_gene (environment.py:542)


@pytest.mark.parametrize("ti", range(4))
def test_non_static_trajectory_matches_reference(ti):
    z, meta = _load()
    m = meta[ti]
    g = lambda k: z[f"t{ti}_{k}"]
    o = OracleWorlds(1, m["height"], m["width"], m["n_genes"], m["max_agents"], seed=m["seed"], world_id0=m["world"],
                     static_families=False)
    o.reset()
    n = int(o.n[0])
    assert (o.type[0] == g("reset_type")).all() and rec_equal(o.rec[0, :n], g("reset_rec"), FIELDS)
    assert o.ns[0].max_gene == m["n_genes"] and _best(o)[0] == [-1 - k for k in range(10)]
    off = dict(a=0, s=0, u=0)
    for t in range(m["steps"]):
        acts = g("actions")[off["a"]:off["a"] + n]
        assert len(acts) == n
        o.rec["action"][0, :n] = acts
        off["a"] += n
        for phase, key in (("step", "s"), ("upd", "u")):
            if phase == "step":
                o.step()
            else:
                o.update()
            n = int(g(phase + "_n")[t])
            assert int(o.n[0]) == n, (t, phase)
            lo = off[key]
            assert (o.type[0] == g(phase + "_type")[t]).all(), (t, phase)
            want = g(phase + "_rec")[lo:lo + n]
            assert rec_equal(o.rec[0, :n], want, FIELDS), (t, phase, rec_diff(o.rec[0, :n], want, FIELDS))
            assert (o.fitness[0, :n].view(np.uint64) == g(phase + "_fitness")[lo:lo + n].view(np.uint64)).all(), (t, phase)
            assert (o.serial[0, :n] == g(phase + "_serial")[lo:lo + n]).all(), (t, phase)
            bs, bf, bb = _best(o)
            assert bs == g(phase + "_best_serial")[t].tolist(), (t, phase)
            assert np.array_equal(np.array(bf).view(np.uint64), g(phase + "_best_fitness")[t].view(np.uint64)), (t, phase)
            assert bb == g(phase + "_best_brain")[t].tolist(), (t, phase)
            if phase == "step":
                assert (o.reward[0, :n].view(np.uint64) == g("step_reward")[lo:lo + n].view(np.uint64)).all(), t
            else:
                assert o.ns[0].max_gene == int(g("upd_max_gene")[t])
                pg, pk = g("upd_produced")[t]
                assert (o.ns[0].produced_gene, o.ns[0].produced_src_best) == (int(pg), int(pk)), t
                if pg >= 0:
                    # the deep-copied brain: best_agents as _update_best_agents left it earlier in the same update_env
                    assert o.ns[0].produced_src_brain == int(g("upd_best_brain")[t][pk]), t
            off[key] += n

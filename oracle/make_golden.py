"""TEST INFRASTRUCTURE.  Mints tests/golden/world_golden.npz by executing the UNMODIFIED reference
(/root/reference, see oracle/ref_harness.py) phase by phase with teacher-forced actions and the
counter-based RNG of include/rl_rng.h.  Run in the build container only:

    python oracle/make_golden.py            # rewrites tests/golden/world_golden.npz

Each case is one phase transition (reset | step | update | topup) of one world:
inputs  = cfg + canonical state before the phase (cell types, agent list with actions),
outputs = canonical state after it + reward (f64) + the 153-float64 observation of every agent.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "world_golden.npz")


class Recorder:
    def __init__(self):
        self.meta, self.arrays = [], {}

    def add(self, phase, cfg, t, world, before, after, extra=None):
        k = len(self.meta)
        m = dict(phase=phase, t=int(t), world=int(world), **cfg)
        if extra:
            m.update(extra)
        self.meta.append(m)
        if before is not None:
            self.arrays[f"c{k}_in_type"] = before["type"]
            self.arrays[f"c{k}_in_rec"] = before["rec"]
        self.arrays[f"c{k}_out_type"] = after["type"]
        self.arrays[f"c{k}_out_rec"] = after["rec"]
        self.arrays[f"c{k}_out_reward"] = after["reward"]
        self.arrays[f"c{k}_out_obs"] = after["obs"]

    def save(self, path):
        np.savez_compressed(path, meta=np.frombuffer(json.dumps(self.meta).encode(), np.uint8), **self.arrays)


def cfg_dict(H, W, G, max_agents, seed, limit_reproduction=False, incentivize_killing=True):
    return dict(height=H, width=W, n_genes=G, max_agents=max_agents, seed=seed,
                limit_reproduction=int(limit_reproduction), incentivize_killing=int(incentivize_killing))


def make_world(cfg, world):
    brains = [rh.NullBrain() for _ in range(cfg["n_genes"])]
    return rh.RefWorld(brains, seed=cfg["seed"], world=world, width=cfg["width"], height=cfg["height"],
                       max_agents=cfg["max_agents"], limit_reproduction=bool(cfg["limit_reproduction"]),
                       incentivize_killing=bool(cfg["incentivize_killing"]))


def with_actions(dump, actions):
    d = dict(dump)
    rec = d["rec"].copy()
    rec["action"] = actions
    d["rec"] = rec
    return d


def trajectory(rec, cfg, world, steps, rng, top_up=None, action_bias=None):
    w = make_world(cfg, world)
    w.reset()
    rec.add("reset", cfg, 0, world, None, w.dump())
    if top_up:
        before = w.dump()
        w.top_up(top_up)
        rec.add("topup", cfg, w.t, world, before, w.dump(), dict(target=top_up))
    for _ in range(steps):
        n = len(w.env.agents)
        if action_bias == "attack":
            actions = rng.choice(8, size=n, p=[.05, .05, .05, .05, .2, .2, .2, .2])
        else:
            actions = rng.integers(0, 8, size=n)
        before = with_actions(w.dump(), actions)
        w.force_actions(actions)
        w.step()
        after = w.dump()
        rec.add("step", cfg, w.t, world, before, after)
        w.update_env(0)
        rec.add("update", cfg, w.t, world, after, w.dump())
        if top_up:
            before = w.dump()
            w.top_up(top_up)
            rec.add("topup", cfg, w.t, world, before, w.dump(), dict(target=top_up))


def crafted(rec, cfg, world, typ, agents, t=1):
    """agents: list of dict(cell, health, age, max_age, gene, flags, action)."""
    w = make_world(cfg, world)
    r = np.zeros(len(agents), rh.REC_DTYPE)
    agents = sorted(agents, key=lambda a: a["cell"])
    for s, a in enumerate(agents):
        r[s] = (a["cell"], a.get("health", 200), a.get("age", 0), a.get("max_age", 50), a.get("gene", 0),
                a.get("flags", 0), a["action"], s)
    w.load_state(typ, r)
    w.observe()
    w.t = t - 1
    before = w.dump()
    assert (before["rec"]["cell"] == r["cell"]).all()
    w.force_actions(before["rec"]["action"])
    w.step()
    after = w.dump()
    rec.add("step", cfg, t, world, before, after)
    w.update_env(0)
    rec.add("update", cfg, t, world, after, w.dump())
    return after


def kats(rec):
    """SURVEY Appendix D known-answer scenarios on an 8x8 grid."""
    cfg = cfg_dict(8, 8, 2, 100, seed=11)
    E = lambda: np.zeros((8, 8), np.uint8)  # noqa: E731
    c = lambda i, j: i * 8 + j              # noqa: E731
    crafted(rec, cfg, 0, E(), [dict(cell=c(2, 2), action=1), dict(cell=c(2, 3), action=1)])                  # D1
    crafted(rec, cfg, 1, E(), [dict(cell=c(2, 2), action=3), dict(cell=c(2, 3), action=3)])                  # D2
    crafted(rec, cfg, 2, E(), [dict(cell=c(2, 2), action=1), dict(cell=c(2, 3), action=3)])                  # D3
    crafted(rec, cfg, 3, E(), [dict(cell=c(2, 2), action=5, gene=0), dict(cell=c(2, 3), action=7, gene=1)])  # D4
    crafted(rec, cfg, 4, E(), [dict(cell=c(2, 2), action=1), dict(cell=c(2, 4), action=3), dict(cell=c(2, 1), action=1)])  # D5
    t = E(); t[2, 3] = 5
    crafted(rec, cfg, 5, t, [dict(cell=c(2, 2), action=1)])                                                  # D6
    t = E(); t[2, 3] = 2
    crafted(rec, cfg, 6, t, [dict(cell=c(2, 2), action=1, health=30)])                                       # D7
    crafted(rec, cfg, 7, E(), [dict(cell=c(0, 0), action=0)])                                                # D8a
    crafted(rec, cfg, 8, E(), [dict(cell=c(0, 0), action=3)])                                                # D8b
    crafted(rec, cfg, 9, E(), [dict(cell=c(2, 2), action=5, gene=0), dict(cell=c(2, 3), action=0, gene=0),
                               dict(cell=c(5, 5), action=0, gene=1)])                                        # D9
    t = E(); t[1, 3] = 1
    crafted(rec, cfg, 10, t, [dict(cell=c(2, 2), action=5), dict(cell=c(2, 3), action=0)])                   # D10
    for k, ma in enumerate([50, 60, 72, 86, 103, 123, 147, 176]):                                            # D11
        t = E(); t[2, 3] = 5
        crafted(rec, cfg, 11 + k, t, [dict(cell=c(2, 2), action=1, max_age=ma, age=3)])
    # agent at (0,0): float health plane (A.8); dead agent's gene plane visible to neighbours
    crafted(rec, cfg, 30, E(), [dict(cell=c(0, 0), action=4, health=120, gene=1), dict(cell=c(0, 1), action=7, health=60, gene=0),
                                dict(cell=c(1, 1), action=0, health=10, gene=1), dict(cell=c(7, 7), action=2, age=49, gene=0)])
    # 3-cycle chain of followers + wrap-around column
    crafted(rec, cfg, 31, E(), [dict(cell=c(4, 5), action=1), dict(cell=c(4, 6), action=1), dict(cell=c(4, 7), action=1),
                                dict(cell=c(4, 0), action=2), dict(cell=c(5, 0), action=3)])


def random_small(rec, rng, n_cases):
    for k in range(n_cases):
        H, W = int(rng.integers(3, 10)), int(rng.integers(3, 10))
        G = int(rng.integers(1, 4))
        cfg = cfg_dict(H, W, G, int(rng.choice([2, 5, 100])), seed=int(rng.integers(1 << 30)),
                       limit_reproduction=bool(rng.integers(2)), incentivize_killing=bool(rng.integers(2)))
        dens = float(rng.choice([0.2, 0.5, 0.8]))
        typ = rng.choice([0, 1, 2, 5], size=(H, W), p=[.7, .15, .1, .05]).astype(np.uint8)
        agents = []
        for cell in range(H * W):
            if rng.random() < dens:
                typ[cell // W, cell % W] = 0
                max_age = int(rng.choice([50, 60, 72]))
                agents.append(dict(cell=cell, health=int(10 * rng.integers(1, 21)),
                                   age=int(rng.choice([0, 3, 7, 20, max_age - 2, max_age - 1])), max_age=max_age,
                                   gene=int(rng.integers(G)), flags=int(rng.choice([0, 8, 16, 24])),
                                   action=int(rng.integers(8))))
        if not agents:
            continue
        crafted(rec, cfg, 100 + k, typ, agents, t=int(rng.integers(1, 50)))


def main():
    rng = np.random.default_rng(20261017)
    rec = Recorder()
    kats(rec)
    random_small(rec, rng, 120)
    trajectory(rec, cfg_dict(30, 30, 2, 100, seed=1), 0, 25, rng)
    trajectory(rec, cfg_dict(30, 30, 3, 100, seed=2, limit_reproduction=True, incentivize_killing=False), 5, 15, rng)
    trajectory(rec, cfg_dict(30, 30, 2, 100, seed=3), 7, 8, rng, top_up=100)
    trajectory(rec, cfg_dict(12, 17, 2, 10, seed=4), 2, 30, rng, top_up=40, action_bias="attack")
    trajectory(rec, cfg_dict(5, 4, 2, 100, seed=5), 1, 20, rng, top_up=12)
    rec.save(OUT)
    print(f"wrote {len(rec.meta)} cases -> {OUT} ({os.path.getsize(OUT) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE.  Mints tests/golden/tracker_golden.npz by executing the UNMODIFIED reference
(/root/reference via oracle/ref_harness.py) with training=True, so that its own Tracker
(ReinLife/Helpers/tracker.py:107-132,178-282) records the per-step `track_results` series and the per-interval
averaged `results`, on teacher-forced trajectories (counter RNG of include/rl_rng.h).  Run in the build container only:

    python oracle/make_tracker_golden.py

Per trajectory: cfg, the forced actions of every step (concatenated, with per-step agent counts), the eight per-step
series (genes x steps; "Avg Number of Populations" is one series), and the averaged results per update interval.
One trajectory is a tiny world that goes extinct for some steps (the reference then appends -1 to EVERY series).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "tracker_golden.npz")
VARS = ["Avg Population Size", "Avg Population Age", "Avg Population Fitness", "Best Population Age",
        "Avg Number of Attacks", "Avg Number of Kills", "Avg Number of Intra Kills"]


def run(cfg, world, steps, interval, rng, p_actions):
    G = cfg["n_genes"]
    brains = [rh.NullBrain() for _ in range(G)]
    w = rh.RefWorld(brains, seed=cfg["seed"], world=world, width=cfg["width"], height=cfg["height"],
                    max_agents=cfg["max_agents"], training=True)
    trk = w.env.tracker
    trk.update_interval = interval
    trk.print_results = False
    w.reset()
    actions, counts = [], []
    series = {v: [[] for _ in range(G)] for v in VARS}
    pops = []
    for n_epi in range(steps + 1):
        n = len(w.env.agents)
        a = rng.choice(8, size=n, p=p_actions)
        actions.append(a.astype(np.int8)); counts.append(n)
        w.force_actions(a)
        w.step()
        before = {v: [len(trk.track_results[v][g]) for g in range(G)] for v in VARS}
        w.update_env(n_epi)
        averaged = n_epi % interval == 0 and n_epi != 0         # _aggregate clears the per-step lists (tracker.py:279-282)
        # the value appended this step: read it before the lists are cleared -> re-derive from a shadow copy
        for v in VARS:
            for g in range(G):
                lst = trk.track_results[v][g]
                series[v][g].append(float(run.last[v][g]) if averaged else float(lst[-1]))
        pops.append(float(run.last_pop) if averaged else float(trk.track_results["Avg Number of Populations"][-1]))
    return dict(actions=np.concatenate(actions) if actions else np.zeros(0, np.int8), counts=np.array(counts, np.int32),
                series=np.array([[series[v][g] for g in range(G)] for v in VARS], np.float64),
                populations=np.array(pops, np.float64),
                results=np.array([[trk.results[v][g] for g in range(G)] for v in VARS], np.float64),
                results_pop=np.array(trk.results["Avg Number of Populations"], np.float64))


def install_shadow():
    """_aggregate deletes the per-step lists at every interval boundary; keep the last appended value of every series by
    wrapping Tracker._track_results (the reference source is untouched)."""
    rh.load_reference()
    mod = sys.modules["ReinLife.Helpers.tracker"]
    orig = mod.Tracker._track_results

    def wrapped(self, agents):
        orig(self, agents)
        run.last = {v: {g: self.track_results[v][g][-1] for g in range(len(self.track_results[v]))} for v in VARS}
        run.last_pop = self.track_results["Avg Number of Populations"][-1]

    mod.Tracker._track_results = wrapped


def main():
    install_shadow()
    rng = np.random.default_rng(20261018)
    uniform = [1 / 8] * 8
    attack = [.05, .05, .05, .05, .2, .2, .2, .2]
    cases = [
        (dict(height=30, width=30, n_genes=3, max_agents=100, seed=6), 0, 90, 7, uniform),
        (dict(height=12, width=12, n_genes=2, max_agents=40, seed=7), 3, 120, 10, attack),
        (dict(height=4, width=4, n_genes=3, max_agents=100, seed=8), 1, 400, 25, attack),      # goes extinct for a while
    ]
    arrays, meta = {}, []
    for k, (cfg, world, steps, interval, p) in enumerate(cases):
        out = run(cfg, world, steps, interval, rng, p)
        meta.append(dict(cfg, world=world, steps=steps, interval=interval,
                         extinct_steps=int((out["populations"] == -1).sum())))
        for name, arr in out.items():
            arrays[f"t{k}_{name}"] = arr
    assert meta[2]["extinct_steps"] > 0, "the tiny world never went extinct; pick another seed"
    np.savez_compressed(OUT, meta=np.frombuffer(json.dumps(meta).encode(), np.uint8), **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", meta)


if __name__ == "__main__":
    main()

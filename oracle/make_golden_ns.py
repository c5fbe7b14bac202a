"""TEST INFRASTRUCTURE.  Mints tests/golden/world_ns_golden.npz: whole trajectories of the UNMODIFIED reference run with
static_families=False (teacher-forced actions, counter RNG of include/rl_rng.h; see oracle/ref_harness.py), i.e. the
World side of the non-static path -- offspring keep the parent's gene/brain (environment.py:506-507), _produce creates
gene max_gene+1 from a deepcopy of random.choice(best_agents).brain (:541-547), _update_best_agents (:728-739) on
Agent.fitness.  Recorded after every step() and every update_env(): cell types, agent list, rewards, fitness, object
identity (serial numbers), the ten best agents, max_gene and the _produce event.

    python oracle/make_golden_ns.py         # rewrites tests/golden/world_ns_golden.npz (build container only)
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "world_ns_golden.npz")
TRAJ = [dict(height=9, width=9, n_genes=2, max_agents=20, seed=1, world=0, steps=160, bias=None),
        dict(height=12, width=10, n_genes=3, max_agents=40, seed=2, world=5, steps=160, bias="attack"),
        dict(height=30, width=30, n_genes=2, max_agents=100, seed=3, world=1, steps=120, bias=None),
        dict(height=5, width=5, n_genes=2, max_agents=30, seed=4, world=2, steps=200, bias="move")]


def snapshot(w, acc, phase, extra=None):
    d, ns = w.dump(with_obs=False), w.ns_state()
    acc[phase + "_n"].append(len(d["rec"]))
    acc[phase + "_type"].append(d["type"].reshape(-1))
    acc[phase + "_rec"].append(d["rec"])
    acc[phase + "_fitness"].append(ns["fitness"]); acc[phase + "_serial"].append(ns["serial"])
    acc[phase + "_best_serial"].append(ns["best_serial"]); acc[phase + "_best_fitness"].append(ns["best_fitness"])
    acc[phase + "_best_brain"].append(ns["best_brain"])
    if phase == "step":
        acc["step_reward"].append(d["reward"])
    else:
        acc["upd_max_gene"].append(ns["max_gene"])
        acc["upd_produced"].append(extra)


def main():
    out, meta = {}, []
    for ti, cfg in enumerate(TRAJ):
        rng = np.random.default_rng(100 + ti)
        brains = [rh.NullBrain() for _ in range(cfg["n_genes"])]
        w = rh.RefWorld(brains, seed=cfg["seed"], world=cfg["world"], width=cfg["width"], height=cfg["height"],
                        max_agents=cfg["max_agents"], static_families=False)
        w.reset()
        acc = {k: [] for k in ("actions", "step_n", "step_type", "step_rec", "step_reward", "step_fitness", "step_serial",
                               "step_best_serial", "step_best_fitness", "step_best_brain", "upd_n", "upd_type", "upd_rec",
                               "upd_fitness", "upd_serial", "upd_best_serial", "upd_best_fitness", "upd_best_brain",
                               "upd_max_gene", "upd_produced")}
        d0, ns0 = w.dump(with_obs=False), w.ns_state()
        out[f"t{ti}_reset_type"], out[f"t{ti}_reset_rec"] = d0["type"].reshape(-1), d0["rec"]
        assert ns0["max_gene"] == cfg["n_genes"] and list(ns0["best_serial"]) == [-1 - k for k in range(10)]
        n_prod = n_repl = 0
        for _ in range(cfg["steps"]):
            n = len(w.env.agents)
            p = {None: None, "attack": [.05, .05, .05, .05, .2, .2, .2, .2], "move": [.22, .22, .22, .22, .03, .03, .03, .03]}[cfg["bias"]]
            actions = rng.choice(8, size=n, p=p)
            acc["actions"].append(actions.astype(np.int8))
            w.force_actions(actions)
            w.step()
            snapshot(w, acc, "step")
            mg, best_before = w.env.max_gene, [b._serial for b in w.env.best_agents]
            rh.CTX.last_choice = -1
            w.update_env()
            produced = (w.env.max_gene, rh.CTX.last_choice) if w.env.max_gene != mg else (-1, -1)
            n_prod += produced[0] >= 0
            n_repl += best_before != [b._serial for b in w.env.best_agents]
            snapshot(w, acc, "upd", produced)
        for k, v in acc.items():
            if k.endswith(("_n", "_max_gene")):
                out[f"t{ti}_{k}"] = np.asarray(v, np.int32)
            elif k == "upd_produced":
                out[f"t{ti}_{k}"] = np.asarray(v, np.int32).reshape(-1, 2)
            elif k.endswith(("_type", "_best_serial", "_best_fitness", "_best_brain")):
                out[f"t{ti}_{k}"] = np.stack(v, 0)
            else:
                out[f"t{ti}_{k}"] = np.concatenate(v, 0) if len(v) else np.zeros(0)
        meta.append(dict(cfg, produced=int(n_prod), best_replaced=int(n_repl), final_max_gene=int(w.env.max_gene)))
        print(meta[-1])
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    main()

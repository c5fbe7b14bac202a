"""Sequential train-event semantics at N = 1 (Helpers/trainer.py:95-96 -> World/entities.py:194-208 ->
Models/PERD3QN.py:91-125): the oracles (C world + explicit-algebra brain) replay the 200-step, 713-train()-event run of
the UNMODIFIED reference recorded in tests/golden/seq_golden.npz (oracle/make_seq_golden.py) -- store order, ring
positions and max-priority rule, the weights / ring each train() sees, priorities written back, Adam, target sync
every soft_update_freq episodes -- and must land on the reference's weights after every recorded optimizer step.
Tolerance (the two fp32 runs differ in summation order only): |w - w_ref| <= 1e-4 on >= 99 % of the elements of every
eval-net tensor and <= 2e-3 (two learning-rate steps) on all of them -- Adam normalises every element's update to ~lr
whatever the gradient's size, so an element whose gradient is pure summation noise (dead ReLU units) can step the other
way; per-event loss rtol 1e-3, priorities atol 1e-3."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seq_golden.npz")
KEYS = ("fc.weight", "fc.bias", "adv_fc1.weight", "adv_fc1.bias", "adv_fc2.weight", "adv_fc2.bias",
        "value_fc1.weight", "value_fc1.bias", "value_fc2.weight", "value_fc2.bias")


def load_seq_golden():
    z = np.load(GOLD)
    meta = json.loads(bytes(z["meta"]).decode())
    ev = z["events"]                              # (step, brain, adam step of that brain), global event order
    by_step = {}
    for i, (step, g, k) in enumerate(ev):
        by_step.setdefault((int(step), int(g)), []).append(i)
    return z, meta, by_step


def assert_weights_close(got, want, what):
    d = np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64))
    assert d.max() <= 2e-3 and (d > 1e-4).mean() <= 0.01, (what, float(d.max()), float((d > 1e-4).mean()))


def sd_of(z, prefix):
    return {k: z[f"{prefix}/{k}"] for k in KEYS}


def test_oracle_replays_reference_sequential_learning():
    from oracle.world_oracle import OracleWorlds
    from oracle import brain_oracle as bo
    z, m, by_step = load_seq_golden()
    G, cap = 2, m["capacity"]
    ow = OracleWorlds(1, m["height"], m["width"], G, max_agents=m["max_agents"], seed=m["seed"], world_id0=m["world"])
    ow.reset(); ow.top_up(m["top_up"])
    ev_w = [sd_of(z, f"w0/{g}") for g in range(G)]
    tg_w = [{k: v.copy() for k, v in ev_w[g].items()} for g in range(G)]
    adam_m = [{k: np.zeros_like(v) for k, v in ev_w[g].items()} for g in range(G)]
    adam_v = [{k: np.zeros_like(v) for k, v in ev_w[g].items()} for g in range(G)]
    steps = [0, 0]
    ring = [dict(items=[None] * cap, prio=np.zeros(cap, np.float32), pos=0, len=0) for _ in range(G)]
    actions, counts = z["actions"], z["counts"]
    pos, checked = 0, 0
    for n_epi in range(m["steps"] + 1):
        n = int(ow.n[0])
        assert n == counts[n_epi], n_epi
        a = np.zeros((1, ow.S), np.int8)
        a[0, :n] = actions[pos:pos + n]; pos += n
        state = ow.obs[0, :n].astype(np.float32).copy()
        ow.set_actions(a)
        ow.step()
        n2 = int(ow.n[0])
        k_ev = [0, 0]
        for s in range(n2):
            r = ow.rec[0, s]
            g, age, dead = int(r["gene"]), int(r["age"]), bool(r["flags"] & 32)
            if age <= 1:
                continue
            rg = ring[g]
            maxp = rg["prio"].max() if rg["len"] else 1.0                        # PERD3QN.py:147
            rg["items"][rg["pos"]] = (state[r["prev_slot"]], int(r["action"]), np.float32(ow.reward[0, s]),
                                      ow.obs[0, s].astype(np.float32), float(dead))
            rg["prio"][rg["pos"]] = maxp
            rg["pos"] = (rg["pos"] + 1) % cap
            rg["len"] = min(cap, rg["len"] + 1)
            if n_epi > m["exploration"]:
                if age % m["train_freq"] == 0 or dead:
                    e = by_step[(n_epi, g)][k_ev[g]]; k_ev[g] += 1
                    idx = z["ev_idx"][e]
                    batch = [rg["items"][i] for i in idx]
                    o = np.stack([b[0] for b in batch]); ac = np.array([b[1] for b in batch])
                    rw = np.array([b[2] for b in batch]); no = np.stack([b[3] for b in batch]); dn = np.array([b[4] for b in batch])
                    grads, loss, prio = bo.dueling_event_grads(ev_w[g], tg_w[g], o, ac, rw, no, dn, m["gamma"])
                    np.testing.assert_allclose(loss, z["ev_loss"][e], rtol=1e-3, atol=1e-5, err_msg=f"event {e}")
                    np.testing.assert_allclose(prio, z["ev_prio"][e], rtol=1e-3, atol=1e-3, err_msg=f"event {e}")
                    for i, p in zip(idx, prio):                                  # PERD3QN.py:177-179 (before the optimizer step)
                        rg["prio"][i] = p
                    steps[g] += 1
                    bo.adam_step(ev_w[g], grads, adam_m[g], adam_v[g], steps[g], m["lr"])
                    assert int(z["events"][e][2]) == steps[g]
                    if [g, steps[g]] in m["snaps"]:
                        for k in KEYS:
                            assert_weights_close(ev_w[g][k], z[f"snap/{g}/{steps[g]}/{k}"], f"brain {g} adam step {steps[g]} {k}")
                        checked += 1
                if n_epi % m["soft_update_freq"] == 0:                           # PERD3QN.py:124-125
                    tg_w[g] = {k: v.copy() for k, v in ev_w[g].items()}
        assert all(k_ev[g] == len(by_step.get((n_epi, g), [])) for g in range(G)), n_epi
        ow.update(); ow.top_up(m["top_up"])
    assert steps == m["adam_steps"] and checked == len(m["snaps"])
    for g in range(G):
        for k in KEYS:
            assert_weights_close(ev_w[g][k], z[f"final/{g}/{k}"], f"final {g} {k}")
            assert_weights_close(tg_w[g][k], z[f"final_target/{g}/{k}"], f"final target {g} {k}")
        assert [ring[g]["pos"], ring[g]["len"]] == z[f"final_pos/{g}"].tolist()
        np.testing.assert_allclose(ring[g]["prio"], z[f"final_prio/{g}"], rtol=1e-3, atol=1e-3)

"""Tracker host logic vs the reference's Tracker (ReinLife/Helpers/tracker.py:178-282) on CPU.

tests/golden/tracker_golden.npz holds the per-step `track_results` series and the per-interval averaged `results` the
UNMODIFIED reference recorded on three teacher-forced trajectories (oracle/make_tracker_golden.py), one of which goes
extinct for 18 steps.  Here the C world oracle replays the trajectories, the stats record of rl_world_stats is restated
in numpy, and reinlife_b200.Helpers.tracker.series_from_record / average_rows must reproduce the reference's numbers
(integer-valued series exactly; "Avg Population Fitness" within 1e-6: the device keeps float32(reward))."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tracker_golden.npz")
N_STATS = 8


def load_tracker_golden():
    z = np.load(GOLD)
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def stats_record(rec, reward, n, G):
    """numpy restatement of k_world_stats for ONE world (csrc/stats_kernels.cu)."""
    out = np.zeros(G * N_STATS + 8)
    r = rec[:n]
    present = 0
    for g in range(G):
        m = r["gene"] == g
        c = int(m.sum())
        if c:
            out[g * N_STATS:g * N_STATS + 7] = [c, r["age"][m].sum(), reward[:n][m].astype(np.float32).astype(np.float64).sum(),
                                                r["age"][m].max(), (r["action"][m] >= 4).sum(), (r["flags"][m] & 1).sum(), 1]
            present += 1
    out[G * N_STATS:G * N_STATS + 3] = [n, 1 if n > 0 else 0, present]
    return out


def check_series(row, z, k, step, G, fit_tol=1e-6):
    ser = z[f"t{k}_series"]          # [7 vars, G, steps]
    for g in range(G):
        for vi in range(7):
            want, got = ser[vi, g, step], row[g][vi]
            if vi == 2 and want != -1:
                assert abs(got - want) <= fit_tol * max(1.0, abs(want)), (k, step, g, vi, got, want)
            else:
                assert got == want, (k, step, g, vi, got, want)
    assert row["populations"] == z[f"t{k}_populations"][step], (k, step)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_tracker_series_match_reference(k):
    from oracle.world_oracle import OracleWorlds
    from reinlife_b200.Helpers.tracker import series_from_record, average_rows, VARIABLES
    z, meta = load_tracker_golden()
    m = meta[k]
    G = m["n_genes"]
    ow = OracleWorlds(1, m["height"], m["width"], G, max_agents=m["max_agents"], seed=m["seed"], world_id0=m["world"])
    ow.reset()
    actions, counts = z[f"t{k}_actions"], z[f"t{k}_counts"]
    pos, rows, res_i = 0, [], 0
    for n_epi in range(m["steps"] + 1):
        n = int(ow.n[0])
        assert n == counts[n_epi], (k, n_epi)
        a = np.zeros((1, ow.S), np.int8)
        a[0, :n] = actions[pos:pos + n]; pos += n
        ow.set_actions(a)
        ow.step()
        row = series_from_record(stats_record(ow.rec[0], ow.reward[0], int(ow.n[0]), G), G)
        check_series(row, z, k, n_epi, G)
        rows.append(row)
        if n_epi % m["interval"] == 0 and n_epi != 0:
            res = average_rows(rows[-m["interval"]:], G)
            want = z[f"t{k}_results"][:, :, res_i]
            for vi, var in enumerate(VARIABLES[:-1]):
                np.testing.assert_allclose(res[var], want[vi], rtol=1e-6, atol=1e-9, equal_nan=True, err_msg=f"{k} {var} {res_i}")
            np.testing.assert_allclose(res[VARIABLES[-1]], z[f"t{k}_results_pop"][res_i], rtol=1e-12, equal_nan=True)
            rows, res_i = [], res_i + 1
        ow.update()
    assert res_i == z[f"t{k}_results"].shape[2] and res_i > 0
    if k == 2:
        assert m["extinct_steps"] > 0

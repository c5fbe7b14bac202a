"""Pins the C oracle (oracle/rl_oracle.c) bit-for-bit against vectors minted from the unmodified reference."""
import numpy as np
import pytest

from golden_util import load_cases, rec_equal, rec_diff
from oracle.world_oracle import OracleWorlds


def run_case(c):
    H, W = c["height"], c["width"]
    o = OracleWorlds(1, H, W, c["n_genes"], c["max_agents"], seed=c["seed"], world_id0=c["world"],
                     limit_reproduction=c["limit_reproduction"], incentivize_killing=c["incentivize_killing"])
    if c["phase"] == "reset":
        o.reset()
    else:
        o.load(0, c["in_type"], c["in_rec"])
        o.t = c["t"]
        if c["phase"] == "step":
            o.t = c["t"] - 1
            o.step()
        elif c["phase"] == "update":
            o.update()
        elif c["phase"] == "topup":
            o.top_up(c["target"])
    return o


def test_golden_has_all_phases():
    phases = {c["phase"] for c in load_cases()}
    assert phases == {"reset", "step", "update", "topup"}
    assert len(load_cases()) > 400


@pytest.mark.parametrize("phase", ["reset", "step", "update", "topup"])
def test_oracle_matches_reference_golden(phase):
    n_checked = 0
    for k, c in enumerate(load_cases()):
        if c["phase"] != phase:
            continue
        o = run_case(c)
        n = len(c["out_rec"])
        assert o.n[0] == n, (k, c["phase"], int(o.n[0]), n)
        assert (o.type[0].reshape(c["height"], c["width"]) == c["out_type"]).all(), (k, phase)
        fields = ("cell", "health", "age", "max_age", "gene", "flags", "action", "prev_slot")
        assert rec_equal(o.rec[0, :n], c["out_rec"], fields), (k, phase, rec_diff(o.rec[0, :n], c["out_rec"], fields))
        if phase == "step":
            assert (o.reward[0, :n] == c["out_reward"]).all(), (k, o.reward[0, :n], c["out_reward"])
        # float64 observations must be identical bit for bit
        assert (o.obs[0, :n].view(np.uint64) == c["out_obs"].view(np.uint64)).all(), (k, phase)
        n_checked += 1
    assert n_checked > 0

"""Pins oracle/brain_oracle.py (explicit fp32 algebra, manual backward + Adam) against vectors produced by the
reference's own nn.Modules / PERD3QNAgent.train() / D3QNAgent.train() (tests/golden/brain_golden.npz)."""
import numpy as np
import pytest

from brain_golden_util import golden, state_dict
from oracle import brain_oracle as bo

FWD_TOL = dict(rtol=1e-4, atol=1e-4)   # |Q| is O(10-50) with the pretrained weights; fp32 summation order differs


def test_forward_oracle_matches_reference_modules():
    z = golden()
    obs = z["obs"]
    np.testing.assert_allclose(bo.dueling_forward(state_dict("perd3qn"), obs, True), z["perd3qn_q_rows"], **FWD_TOL)
    np.testing.assert_allclose(bo.dueling_forward(state_dict("perd3qn"), obs[:64], False), z["perd3qn_q_batch64"], **FWD_TOL)
    np.testing.assert_allclose(bo.dueling_forward(state_dict("d3qn"), obs, True), z["d3qn_q_rows"], **FWD_TOL)
    np.testing.assert_allclose(bo.dqn_forward(state_dict("dqn"), obs), z["dqn_q_rows"], **FWD_TOL)
    pi, v = bo.ppo_forward(state_dict("ppo"), obs)
    np.testing.assert_allclose(pi, z["ppo_pi_rows"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(v, z["ppo_v_rows"], **FWD_TOL)


@pytest.mark.parametrize("name", ["perd3qn", "d3qn"])
def test_train_event_oracle_matches_reference_train(name):
    z = golden()
    w = {k: v.copy() for k, v in state_dict(f"train_{name}/w0").items()}
    target = state_dict(f"train_{name}/target")
    m = {k: np.zeros_like(v) for k, v in w.items()}
    v2 = {k: np.zeros_like(v) for k, v in w.items()}
    for step in range(3):
        p = f"train_{name}/s{step}/"
        ev = (z[p + "obs"], z[p + "action"], z[p + "reward"], z[p + "next_obs"], z[p + "done"])
        grads, loss, prio = bo.dueling_event_grads(w, target, *ev, gamma=0.99)
        if name == "perd3qn":   # update_priorities (PERD3QN.py:110-111, 177-179): sequential, last duplicate wins
            expect = z[p + "prio_before"].copy()
            for i, pr in zip(z[p + "indices"], prio):
                expect[i] = pr
            np.testing.assert_allclose(expect, z[p + "prio_after"], rtol=2e-5, atol=2e-5)
        bo.adam_step(w, grads, m, v2, step + 1, lr=1e-3)
        ref_w = state_dict(p + "w")
        for k in w:
            np.testing.assert_allclose(w[k], ref_w[k], rtol=0, atol=3e-6, err_msg=f"{name} step {step} {k}")


def test_sampler_is_proportional_and_exact():
    rng = np.random.default_rng(0)
    prio = rng.random(1000).astype(np.float32) * 5
    w = bo.per_weight(prio)
    u = [int(x) for x in rng.integers(0, 1 << 53, size=20000)]
    idx = np.array(bo.per_sample(w, u))
    # agrees with the float64 cdf / searchsorted(side='right') of np.random.choice except on boundary slivers
    cdf = np.cumsum(w.astype(np.float64)); cdf /= cdf[-1]
    ref = np.searchsorted(cdf, np.array(u, np.float64) / 2.0 ** 53, side="right")
    assert (idx != ref).mean() < 1e-3
    hist = np.bincount(idx, minlength=1000) / len(idx)
    assert abs(hist - w / w.sum()).max() < 5e-3
    assert bo.per_sample(np.zeros(4, np.float32), [5]) == [0]

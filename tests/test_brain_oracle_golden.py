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


def test_uniform_sample_oracle_is_cpython_random_sample():
    """oracle uniform_sample == the stdlib's Random.sample driven by the same _randbelow stream (both branches of the
    algorithm: pool method up to n = 277, set method with redraws above), which is what D3QN.py:140 / DQN.py:100 call."""
    import random

    class Counter(random.Random):
        def __init__(self, bits_fn):
            super().__init__(0)
            self.f, self.c = bits_fn, 0

        def _randbelow(self, n):
            v = ((self.f(self.c) >> 32) * n) >> 32
            self.c += 1
            return v

    from oracle.ref_harness import mix64
    for seed in range(3):
        f = lambda c, s=seed: mix64(0x1234567 * (s + 1) + c)       # noqa: E731
        for n in (64, 65, 100, 276, 277, 278, 300, 1001, 10000):
            for k in (64, 32):
                got = bo.uniform_sample(n, k, f)
                assert got == Counter(f).sample(range(n), k), (n, k)
                assert len(set(got)) == k and min(got) >= 0 and max(got) < n
    with pytest.raises(ValueError):
        bo.uniform_sample(63, 64, f)


def _golden2():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "brain_golden2.npz"))


def _sd2(z, prefix):
    pre = prefix + "/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre) and "/" not in k[len(pre):]}


def test_dqn_train_oracle_matches_reference_train():
    """Five iterations of train() (Models/DQN.py:142-153): explicit smooth-L1 backward + Adam vs the weights the
    reference's optimizer produced after every step (oracle/make_brain_golden2.py)."""
    z = _golden2()
    w = {k: v.copy() for k, v in _sd2(z, "train_dqn/w0").items()}
    target = _sd2(z, "train_dqn/target")
    m = {k: np.zeros_like(v) for k, v in w.items()}
    v2 = {k: np.zeros_like(v) for k, v in w.items()}
    for it in range(5):
        p = f"train_dqn/i{it}/"
        grads, _ = bo.dqn_iter_grads(w, target, z[p + "obs"], z[p + "action"], z[p + "reward"], z[p + "next_obs"], z[p + "done_mask"])
        bo.adam_step(w, grads, m, v2, it + 1, lr=5e-4)
        ref_w = _sd2(z, p + "w")
        for k in w:
            np.testing.assert_allclose(w[k], ref_w[k], rtol=0, atol=3e-6, err_msg=f"iter {it} {k}")


@pytest.mark.parametrize("T", [1, 5, 37, 100])
def test_ppo_learn_oracle_matches_reference_learn(T):
    """Three epochs of PPO.learn() (Models/PPO.py:136-162) on a data list of T transitions: GAE in float32 (NEP 50),
    clipped surrogate + scalar smooth-L1 value loss, explicit backward + Adam vs the reference's weights per epoch."""
    z = _golden2()
    w = {k: v.copy() for k, v in _sd2(z, "train_ppo/w0").items()}
    m = {k: np.zeros_like(v) for k, v in w.items()}
    v2 = {k: np.zeros_like(v) for k, v in w.items()}
    p = f"train_ppo/T{T}/"
    for e in range(3):
        grads, _ = bo.ppo_epoch_grads(w, z[p + "obs"], z[p + "action"], z[p + "reward"], z[p + "next_obs"], z[p + "prob_a"], z[p + "done"])
        bo.adam_step(w, grads, m, v2, e + 1, lr=5e-4)
        ref_w = _sd2(z, f"{p}e{e}")
        for k in w:
            np.testing.assert_allclose(w[k], ref_w[k], rtol=0, atol=2e-5, err_msg=f"T {T} epoch {e} {k}")   # Adam step 1 = lr*g/(|g|+1e-8): near-zero gradients amplify fp32 summation order

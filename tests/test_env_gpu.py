"""End-to-end through the public API (trainer / tester / Environment) on the GPU."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_trainer_runs_and_learns_perd3qn(tmp_path, monkeypatch):
    import reinlife_b200 as rl
    from reinlife_b200.Models import PERD3QN
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    brains = [PERD3QN(exploration=2, train_freq=5, capacity=400), PERD3QN(exploration=2, train_freq=5, capacity=400)]
    w_before = [b.eval_net.state_dict()["fc.weight"].clone() for b in brains]
    env = rl.trainer(brains, n_episodes=30, width=12, height=12, max_agents=30, update_interval=10, print_results=False,
                     save=True, n_worlds=16, seed=3, saturate_to=30)
    torch.cuda.synchronize()
    assert int(env.world.status.max()) == 0
    for b, w0 in zip(brains, w_before):
        assert int(b._dev.adam_step) > 0
        assert not torch.equal(b.eval_net.state_dict()["fc.weight"], w0)
        assert torch.isfinite(b._dev.params).all()
    eps = env.epsilons()
    assert all(abs(e - 0.9 * 0.99 ** 30) < 1e-12 for e in eps)          # n_epi 1..30 decayed once each (PERD3QN.py:82-86)
    # tracker: 3 aggregation points (n_epi = 10, 20, 30), pooled over the 16 worlds
    res = env.tracker.results
    assert len(res["Avg Population Size"][0]) == 3 and len(res["Avg Number of Populations"]) == 3
    assert 10 < res["Avg Population Size"][0][-1] + res["Avg Population Size"][1][-1] <= 31
    # checkpoints in the reference layout, loadable as reference-shaped state_dicts
    import glob
    files = sorted(glob.glob(str(tmp_path / "experiments" / "*" / "PERD3QN" / "brain_gene_*.pt")))
    assert len(files) == 2
    sd = torch.load(files[0])
    assert sd["fc.weight"].shape == (128, 153) and sd["value_fc2.weight"].shape == (1, 128)
    # agents view
    ag = env.agents
    assert len(ag) == int(env.world.n_agents[0]) and all(0 <= a.gene < 2 for a in ag)


def test_tester_inference_dqn_pretrained_weights():
    import reinlife_b200 as rl
    from reinlife_b200.Models import DQN
    from brain_golden_util import state_dict
    b = DQN(training=False)
    b.agent.load_state_dict(state_dict("dqn"))
    env = rl.tester([b], n_worlds=32, n_steps=20, saturate_to=100, seed=1)
    torch.cuda.synchronize()
    assert env.epsilons() == [0.0]
    assert int(env.world.n_agents.min()) == 100
    # plugin call for a single host observation agrees with the batched kernel
    from brain_golden_util import golden
    z = golden()
    a = b.get_action(z["obs"][0], 0)
    assert a == int(np.argmax(z["dqn_q_rows"][0]))


def test_unimplemented_paths_fail_loudly():
    import reinlife_b200 as rl
    from reinlife_b200.Models import DQN
    with pytest.raises(ZeroDivisionError):
        rl.trainer([DQN()], n_episodes=1, save=False, n_worlds=2)           # DQN(max_epi=0) while training (DQN.py:69)
    env = rl.Environment(brains=[DQN(training=False)], training=False, n_worlds=1)
    with pytest.raises(NotImplementedError):
        env.render()


def test_trainer_learns_d3qn_and_reports_short_buffer():
    """D3QN (uniform random.sample replay, D3QN.py:138-142) trains on the device; a train event on a buffer shorter than a
    batch surfaces as the reference's ValueError (raised at the next status check instead of inside the loop)."""
    import reinlife_b200 as rl
    from reinlife_b200.Models import D3QN
    torch.manual_seed(1)
    brains = [D3QN(exploration=12, train_freq=5, capacity=300), D3QN(exploration=12, train_freq=5, capacity=300)]
    w_before = [b.eval_net.state_dict()["fc.weight"].clone() for b in brains]
    env = rl.trainer(brains, n_episodes=40, width=12, height=12, max_agents=30, update_interval=10, print_results=False,
                     save=False, n_worlds=8, seed=5, saturate_to=30, precision="fp32")
    torch.cuda.synchronize()
    for b, w0 in zip(brains, w_before):
        assert int(b._dev.adam_step) > 0 and torch.isfinite(b._dev.params).all()
        assert not torch.equal(b.eval_net.state_dict()["fc.weight"], w0)
        assert int(b._replay.len.max()) == 300                                   # the deque is full and wrapping
    with pytest.raises(ValueError):                                              # ~15 agents/gene x 1 step < 64 items
        rl.trainer([D3QN(exploration=0, train_freq=2), D3QN(exploration=0, train_freq=2)], n_episodes=6, width=12, height=12,
                   max_agents=30, print_results=False, save=False, n_worlds=4, seed=5, saturate_to=30, precision="fp32")


def test_trainer_learns_dqn():
    """DQN trains on the device: 5 Adam steps per step with a train trigger once a ring holds > 1000 items (DQN.py:79),
    target <- agent after every trigger (DQN.py:81), linear epsilon (DQN.py:67-69)."""
    import reinlife_b200 as rl
    from reinlife_b200.Models import DQN
    torch.manual_seed(2)
    b = DQN(max_epi=200, train_freq=5, buffer_limit=1500)
    w0 = b.agent.state_dict()["fc1.weight"].clone()
    env = rl.trainer([b], n_episodes=60, width=12, height=12, max_agents=40, update_interval=20, print_results=False,
                     save=False, n_worlds=6, seed=8, saturate_to=40)
    torch.cuda.synchronize()
    assert int(b._replay.len.max()) == 1500
    steps = int(b._dev.adam_step)
    assert steps > 0 and steps % 5 == 0
    assert torch.isfinite(b._dev.params).all() and not torch.equal(b.agent.state_dict()["fc1.weight"], w0)
    for k, v in b.agent.state_dict().items():                       # target == agent after the last trigger
        assert torch.equal(v, b.target.state_dict()[k])
    assert abs(env.epsilons()[0] - max(0.01, 0.20 - 0.20 * (60 / 200))) < 1e-12


def test_trainer_learns_ppo_with_perd3qn():
    """PPO trains on the device next to a PERD3QN brain (BASELINE.json configs[4] brain mix, static families)."""
    import reinlife_b200 as rl
    from reinlife_b200.Models import PPO, PERD3QN
    torch.manual_seed(3)
    brains = [PPO(train_freq=5), PERD3QN(exploration=2, train_freq=5, capacity=400)]
    w0 = brains[0].model.state_dict()["fc1.weight"].clone()
    env = rl.trainer(brains, n_episodes=25, width=12, height=12, max_agents=30, update_interval=10, print_results=False,
                     save=False, n_worlds=8, seed=6, saturate_to=30)
    torch.cuda.synchronize()
    b = brains[0]
    steps = int(b._dev.adam_step)
    assert steps > 0 and steps % 3 == 0                                   # k_epoch optimizer steps per learn step
    assert torch.isfinite(b._dev.params).all() and not torch.equal(b.model.state_dict()["fc1.weight"], w0)
    assert int(b._replay.status) == 0 and int(b._replay.traj.len.max()) < 200
    assert int(brains[1]._dev.adam_step) > 0


@pytest.mark.parametrize("k", [0, 1, 2])
def test_tracker_matches_reference_tracker(k):
    """(f)1: Environment(training=True).update_env -> rl_world_stats -> Tracker vs the per-step `track_results` series and
    the per-interval averaged `results` recorded from the UNMODIFIED reference's Tracker (Helpers/tracker.py:178-282) on
    teacher-forced trajectories (tests/golden/tracker_golden.npz); trajectory 2 goes extinct for 18 steps (all -1)."""
    import reinlife_b200 as rl
    from reinlife_b200.Models import PERD3QN
    from reinlife_b200.Helpers.tracker import VARIABLES
    from test_tracker_cpu import load_tracker_golden, check_series
    z, meta = load_tracker_golden()
    m = meta[k]
    G = m["n_genes"]
    brains = [PERD3QN(capacity=64) for _ in range(G)]
    env = rl.Environment(width=m["width"], height=m["height"], brains=brains, max_agents=m["max_agents"],
                         update_interval=m["interval"], print_results=False, training=True, n_worlds=1, seed=m["seed"],
                         world_id0=m["world"])
    env.tracker.history = []
    env.reset()
    actions, counts = z[f"t{k}_actions"], z[f"t{k}_counts"]
    pos = 0
    for n_epi in range(m["steps"] + 1):
        n = int(env.world.n_agents[0])
        assert n == counts[n_epi], (k, n_epi)
        a = np.zeros((1, env.world.S), np.int8)
        a[0, :n] = actions[pos:pos + n]; pos += n
        env.world.set_actions(a)
        env.step()
        env.update_env(n_epi)
    env.tracker._drain(env.tracker.k % env.tracker.ring_len)
    hist = env.tracker.history
    assert len(hist) == m["steps"] + 1
    for step, row in enumerate(hist):
        check_series(row, z, k, step, G)
    want = z[f"t{k}_results"]
    for vi, var in enumerate(VARIABLES[:-1]):
        for g in range(G):
            np.testing.assert_allclose(env.tracker.results[var][g], want[vi, g], rtol=1e-6, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(env.tracker.results[VARIABLES[-1]], z[f"t{k}_results_pop"], rtol=1e-12, equal_nan=True)


def test_saved_parameters_carry_the_decayed_epsilon(tmp_path, monkeypatch):
    """parameters_gene_*.json and brain.epsilon / brain.n_epi follow the device schedule (the reference writes the decayed
    values, e.g. pretrained/PERD3QN/.../parameters_gene_0.json)."""
    import glob
    import json
    import reinlife_b200 as rl
    from reinlife_b200.Models import PERD3QN
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    brains = [PERD3QN(exploration=1000, capacity=64), PERD3QN(exploration=1000, capacity=64)]
    env = rl.trainer(brains, n_episodes=25, width=10, height=10, max_agents=20, update_interval=10, print_results=False,
                     save=True, n_worlds=4, seed=2, saturate_to=20)
    eps = env.epsilons()
    for g, b in enumerate(brains):
        assert b.epsilon == eps[g] and abs(b.epsilon - 0.9 * 0.99 ** 25) < 1e-12 and b.n_epi == 25
    for f in sorted(glob.glob(str(tmp_path / "experiments" / "*" / "PERD3QN" / "parameters_gene_*.json"))):
        p = json.load(open(f))
        assert abs(p["epsilon"] - 0.9 * 0.99 ** 25) < 1e-12 and p["n_epi"] == 25

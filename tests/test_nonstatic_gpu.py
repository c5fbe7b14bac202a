"""static_families=False end to end (World/environment.py:149,506-507,541-547,728-739): device World kernels + per-world
brain pools of plugin-mode brains (reinlife_b200/World/nonstatic.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_brain_clone_is_a_deep_copy():
    """copy.deepcopy(brain) semantics: weights, target, Adam state, replay ring and schedule scalars are copied; the
    copy and the original evolve independently afterwards."""
    from reinlife_b200.Models import PERD3QN
    from brain_golden_util import golden
    obs = golden()["obs"]
    rng = np.random.default_rng(0)
    torch.manual_seed(0)
    b = PERD3QN(exploration=0, train_freq=4, capacity=128)

    def feed(brain, k, age0):
        for t in range(k):
            brain.learn(age=age0 + t, dead=False, action=int(rng.integers(8)), state=obs[rng.integers(512)], reward=0.3,
                        state_prime=obs[rng.integers(512)], done=False, n_epi=1)
    feed(b, 70, 2)                                       # 70 stores; train triggers once 64 are there... (ages 68 = 17 * 4)
    b.epsilon, b.n_epi = 0.5, 7
    c = b.clone()
    assert c is not b and c._dev is not b._dev and c._replay is not b._replay and c._plugin_host is not b._plugin_host
    assert (c.epsilon, c.n_epi, c.train_freq, c.capacity) == (0.5, 7, 4, 128)
    for name in ("params", "target", "adam_m", "adam_v", "adam_step"):
        assert torch.equal(getattr(c._dev, name), getattr(b._dev, name)), name
    for name in ("obs", "next_obs", "action", "reward", "done", "prio", "pw", "len", "pos"):
        assert torch.equal(getattr(c._replay, name), getattr(b._replay, name)), name
    before = b._dev.params.clone()
    steps_b = int(b._dev.adam_step)
    feed(c, 8, 72)                                       # the copy trains on ...
    assert int(c._dev.adam_step) > steps_b and int(b._dev.adam_step) == steps_b
    assert torch.equal(b._dev.params, before) and not torch.equal(c._dev.params, before)
    assert int(c._replay.len[0]) == int(b._replay.len[0]) + 8


def test_trainer_non_static_families_runs_and_evolves():
    import reinlife_b200 as rl
    from reinlife_b200.Models import PERD3QN, PPO
    torch.manual_seed(1)
    brains = [PPO(train_freq=5), PERD3QN(exploration=5, train_freq=4, capacity=256)]
    env = rl.trainer(brains, n_episodes=160, width=12, height=12, max_agents=30, update_interval=40, print_results=False,
                     static_families=False, save=False, n_worlds=2, seed=3)
    assert type(env).__name__ == "NonStaticEnvironment" and env.static_families is False
    env.check_status()
    states = env.world.ns_host()
    n = env.world.n_agents.cpu().numpy()
    rec = env.world.rec_host()
    assert env.max_gene > 2 and env.max_gene == max(s.max_gene for s in states)      # _produce created new lineages
    produced = sum(s.max_gene - 2 for s in states)
    assert produced >= 5
    learned = 0
    for w in range(2):
        live = set(int(g) for g in rec[w, :n[w]]["gene"])
        keep = live | set(b.brain for b in states[w].best if b.brain >= 0)
        assert set(env.lineages(w)) == keep or set(env.lineages(w)) >= live             # every live lineage has its brain
        for g in env.lineages(w):
            b = env.pools[w][g]
            assert b.method in ("PPO", "PERD3QN")
            if b._dev is not None and b._plugin_host is not None:
                learned += int(b._dev.adam_step) > 0
                assert torch.isfinite(b._dev.params).all()
    assert learned > 0
    res = env.tracker.results
    assert len(res["Avg Population Size"][0]) == 4 and len(res["Avg Number of Populations"]) == 4
    assert len(env.best_agents) == 10 and all(hasattr(a.brain, "method") for a in env.best_agents)
    # worlds are independent runs with their own brain objects: world 1 never sees world 0's objects
    assert not (set(map(id, env.pools[0].values())) & set(map(id, env.pools[1].values())))


def test_non_static_checkpoint_layout(tmp_path, monkeypatch):
    """environment.py:233-256 with families=False: best agents' brains saved as brain_<k>.pt per method."""
    import glob
    import reinlife_b200 as rl
    from reinlife_b200.Models import PERD3QN
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(2)
    env = rl.trainer([PERD3QN(exploration=1000), PERD3QN(exploration=1000)], n_episodes=30, width=10, height=10, max_agents=20,
                     update_interval=10, print_results=False, static_families=False, save=True, n_worlds=1, seed=1)
    files = sorted(glob.glob(str(tmp_path / "experiments" / "*" / "PERD3QN" / "brain_*.pt")))
    assert len(files) == 10
    sd = torch.load(files[0])
    assert sd["fc.weight"].shape == (128, 153)

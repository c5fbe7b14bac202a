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
        # (raw bytes: the rows of a ring are allocated uninitialised, unfilled slots may hold NaN bit patterns)
        x, y = getattr(c._replay, name), getattr(b._replay, name)
        assert x.dtype == y.dtype and torch.equal(x.contiguous().view(torch.uint8), y.contiguous().view(torch.uint8)), name
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


def test_non_static_learning_run_matches_reference():
    """260 steps of ONE reference world with static_families=False and learning on (tests/golden/ns_learn_golden.npz,
    oracle/make_ns_learn_golden.py: 336 train() events, 15 _produce deep copies incl. replay memory and Adam state,
    mutate_brain, best-agent bookkeeping) replayed through NonStaticEnvironment: the same lineage triggers every event,
    per-event loss, and at the end every referenced brain's weights, target, Adam step count and ring fill."""
    import json
    import os
    import reinlife_b200 as rl
    from reinlife_b200 import plugin
    from reinlife_b200.Models import PERD3QN
    from test_seq_cpu import KEYS, assert_weights_close
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ns_learn_golden.npz"))
    m = json.loads(bytes(z["meta"]).decode())
    brains = []
    for g in range(2):
        b = PERD3QN(exploration=m["exploration"], train_freq=m["train_freq"], capacity=m["capacity"],
                    soft_update_freq=m["soft_update_freq"], learning_rate=m["lr"], gamma=m["gamma"])
        sd = {k: z[f"w0/{g}/{k}"] for k in KEYS}
        b.eval_net.load_state_dict(sd); b.target_net.load_state_dict(sd)
        brains.append(b)
    env = rl.Environment(width=m["width"], height=m["height"], brains=brains, max_agents=m["max_agents"], print_results=False,
                         static_families=False, training=False, n_worlds=1, seed=m["seed"], world_id0=m["world"])
    env.training = True            # (training=False above only keeps the tracker out, like the minting run)
    cur = {"e": 0, "n_epi": 0}
    gene_of = lambda brain: next(g for g, b in env.pools[0].items() if b is brain)   # noqa: E731

    def override(brain):
        e = cur["e"]
        assert int(z["ev_gene"][e]) == gene_of(brain) and int(z["ev_step"][e]) == cur["n_epi"], (e, gene_of(brain))
        return z["ev_idx"][e]

    def hook(brain):
        e = cur["e"]
        np.testing.assert_allclose(float(brain._dev.loss[0]), z["ev_loss"][e], rtol=2e-3, atol=1e-5, err_msg=f"event {e}")
        cur["e"] += 1
    plugin.SAMPLE_OVERRIDE, plugin.EVENT_HOOK = override, hook
    try:
        env.reset()
        pos, produced = 0, []
        for n_epi in range(m["steps"] + 1):
            cur["n_epi"] = n_epi
            n, _ = env.snapshot_state()
            assert int(n[0]) == z["counts"][n_epi], n_epi
            a = np.zeros((1, env.world.S), np.int8)
            a[0, :n[0]] = z["actions"][pos:pos + n[0]]; pos += int(n[0])
            env.world.set_actions(a)
            env.step()
            env.learn(n_epi)
            mg = env.max_gene
            env.tracker.update_interval = 10 ** 9
            env.update_env(n_epi)
            if env.max_gene > mg:
                produced.append((n_epi, env.max_gene, env.world.ns_host()[0].produced_src_best))
    finally:
        plugin.SAMPLE_OVERRIDE = plugin.EVENT_HOOK = None
    assert cur["e"] == m["n_events"] and env.max_gene == m["max_gene"]
    assert produced == [tuple(int(x) for x in r) for r in z["produced"]]
    for g, steps, ring_len, ring_pos in m["final"]:
        b = env.pools[0][g]
        assert (int(b._dev.adam_step), int(b._replay.len[0]), int(b._replay.pos[0])) == (steps, ring_len, ring_pos), g
        got, tgt = b.eval_net.state_dict(), b.target_net.state_dict()
        for k in KEYS:
            assert_weights_close(got[k].numpy(), z[f"final/{g}/{k}"], f"final gene {g} {k}")
            assert_weights_close(tgt[k].numpy(), z[f"final_target/{g}/{k}"], f"final target gene {g} {k}")

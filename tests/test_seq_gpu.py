"""The reference's per-agent order of learning effects on the device.

* Environment(sequential_events=True): the 200-step / 713-train()-event single-world run of the UNMODIFIED reference
  (tests/golden/seq_golden.npz, see tests/test_seq_cpu.py for what is recorded) replayed through the public Environment
  API on the fp32 path: per-event loss and priorities, the eval-net weights after every recorded optimizer step, the
  target nets, the rings' write positions and priorities at the end.
* brain.learn(...) with HOST buffers (the ReinLife.Models plugin surface, World/entities.py:194-208): the same run
  driven agent by agent through PERD3QN.learn from a CPU world (the C oracle) -- a stand-in for the reference's own
  Environment/Agent classes calling our brains -- and PPO.learn against the call-by-call recording of the
  reference's PPOAgent.learn (tests/golden/plugin_golden.npz); DQN / PERDQN / D3QN learn() behaviour (buffer thresholds,
  optimizer-step counts, target sync, epsilon, the D3QN ValueError).
Tolerances: see tests/test_seq_cpu.py (weights: 1e-4 on >= 99 % of the elements, 2e-3 everywhere)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_seq_cpu import KEYS, assert_weights_close, load_seq_golden, sd_of   # noqa: E402

pytestmark = pytest.mark.gpu


def _brains(z, m):
    from reinlife_b200.Models import PERD3QN
    brains = []
    for g in range(2):
        b = PERD3QN(exploration=m["exploration"], train_freq=m["train_freq"], capacity=m["capacity"],
                    soft_update_freq=m["soft_update_freq"], learning_rate=m["lr"], gamma=m["gamma"])
        b.eval_net.load_state_dict(sd_of(z, f"w0/{g}"))
        b.target_net.load_state_dict(sd_of(z, f"w0/{g}"))
        brains.append(b)
    return brains


def _check_final(z, brains):
    for g, b in enumerate(brains):
        got, tgt = b.eval_net.state_dict(), b.target_net.state_dict()
        for k in KEYS:
            assert_weights_close(got[k].numpy(), z[f"final/{g}/{k}"], f"final {g} {k}")
            assert_weights_close(tgt[k].numpy(), z[f"final_target/{g}/{k}"], f"final target {g} {k}")
        assert [int(b._replay.pos[0]), int(b._replay.len[0])] == z[f"final_pos/{g}"].tolist()
        np.testing.assert_allclose(b._replay.prio[0].cpu().numpy(), z[f"final_prio/{g}"], rtol=1e-3, atol=1e-3)


def test_sequential_events_match_reference_run():
    import reinlife_b200 as rl
    z, m, by_step = load_seq_golden()
    brains = _brains(z, m)
    env = rl.Environment(width=m["width"], height=m["height"], brains=brains, max_agents=m["max_agents"], print_results=False,
                         training=True, n_worlds=1, seed=m["seed"], world_id0=m["world"], precision="fp32",
                         sequential_events=True, update_interval=10 ** 9)
    cur = {"n_epi": 0, "seen": 0, "snaps": 0}

    def override(g, k):
        return z["ev_idx"][by_step[(cur["n_epi"], g)][k]]

    def hook(g, k):
        e = by_step[(cur["n_epi"], g)][k]
        dev = brains[g]._dev
        np.testing.assert_allclose(float(dev.loss[0]), z["ev_loss"][e], rtol=1e-3, atol=1e-5, err_msg=f"event {e}")
        np.testing.assert_allclose(dev.new_prio[0].cpu().numpy(), z["ev_prio"][e], rtol=1e-3, atol=1e-3, err_msg=f"event {e}")
        step = int(dev.adam_step)
        assert step == int(z["events"][e][2])
        if [g, step] in m["snaps"]:
            got = brains[g].eval_net.state_dict()
            for key in KEYS:
                assert_weights_close(got[key].numpy(), z[f"snap/{g}/{step}/{key}"], f"brain {g} adam step {step} {key}")
            cur["snaps"] += 1
        cur["seen"] += 1

    env.sample_override, env.event_hook = override, hook
    env.reset(); env.top_up(m["top_up"])
    actions, counts = z["actions"], z["counts"]
    pos = 0
    for n_epi in range(m["steps"] + 1):
        cur["n_epi"] = n_epi
        n = int(env.world.n_agents[0])
        assert n == counts[n_epi], n_epi
        a = np.zeros((1, env.world.S), np.int8)
        a[0, :n] = actions[pos:pos + n]; pos += n
        env.world.set_actions(a)
        env.step()
        env.learn(n_epi)
        env.update_env(n_epi)
        env.top_up(m["top_up"])
    assert cur["seen"] == len(z["events"]) and cur["snaps"] == len(m["snaps"])
    assert [int(b._dev.adam_step) for b in brains] == m["adam_steps"]
    _check_final(z, brains)


def test_plugin_learn_perd3qn_driven_agent_by_agent():
    """The loop of Helpers/trainer.py:95-96 with a HOST world: every agent's transition goes through brain.learn(...)
    with numpy observations, like Agent.learn (World/entities.py:194-208) calls it."""
    from oracle.world_oracle import OracleWorlds
    from reinlife_b200.plugin import PluginHost
    z, m, by_step = load_seq_golden()
    brains = _brains(z, m)
    cur = {"n_epi": 0, "k": [0, 0], "seen": 0}
    for g, b in enumerate(brains):
        b._plugin_host = PluginHost(b)
        b._plugin_host.env.sample_override = (lambda _g, _k, g=g: z["ev_idx"][by_step[(cur["n_epi"], g)][cur["k"][g]]])

        def hook(_g, _k, g=g, b=b):
            e = by_step[(cur["n_epi"], g)][cur["k"][g]]
            np.testing.assert_allclose(float(b._dev.loss[0]), z["ev_loss"][e], rtol=1e-3, atol=1e-5, err_msg=f"event {e}")
            cur["k"][g] += 1; cur["seen"] += 1
        b._plugin_host.env.event_hook = hook
    ow = OracleWorlds(1, m["height"], m["width"], 2, max_agents=m["max_agents"], seed=m["seed"], world_id0=m["world"])
    ow.reset(); ow.top_up(m["top_up"])
    actions = z["actions"]
    pos = 0
    for n_epi in range(m["steps"] + 1):
        cur["n_epi"], cur["k"] = n_epi, [0, 0]
        n = int(ow.n[0])
        a = np.zeros((1, ow.S), np.int8)
        a[0, :n] = actions[pos:pos + n]; pos += n
        state = ow.obs[0, :n].copy()
        ow.set_actions(a)
        ow.step()
        for s in range(int(ow.n[0])):                       # for agent in env.agents: agent.learn(n_epi=n_epi)
            r = ow.rec[0, s]
            if int(r["age"]) > 1:                           # World/entities.py:196
                dead = bool(r["flags"] & 32)
                brains[int(r["gene"])].learn(age=int(r["age"]), dead=dead, action=int(r["action"]), state=state[r["prev_slot"]],
                                             reward=float(ow.reward[0, s]), state_prime=ow.obs[0, s], done=dead, n_epi=n_epi)
        ow.update(); ow.top_up(m["top_up"])
    assert cur["seen"] == len(z["events"])
    assert [int(b._dev.adam_step) for b in brains] == m["adam_steps"]
    _check_final(z, brains)


def test_plugin_learn_ppo_matches_reference_calls():
    from reinlife_b200.Models import PPO
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(path, "plugin_golden.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    obs = np.load(os.path.join(path, "brain_golden.npz"))["obs"]
    keys = ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc_pi.weight", "fc_pi.bias", "fc_v.weight", "fc_v.bias")
    b = PPO(train_freq=meta["train_freq"])
    b.model.load_state_dict({k: z[f"w0/{k}"] for k in keys})
    n_learn = 0
    for c, prob in zip(z["calls"], z["probs"]):
        before = int(b._dev.adam_step) if b._dev is not None and b._plugin_host is not None else 0
        b.learn(age=int(c["age"]), dead=bool(c["dead"]), action=int(c["action"]), state=obs[c["s"]], reward=float(c["reward"]),
                state_prime=obs[c["sp"]], done=bool(c["dead"]), prob=torch.from_numpy(prob))
        after = int(b._dev.adam_step)
        assert after - before in (0, 3)                                           # k_epoch optimizer steps per learn()
        if after != before:
            n_learn += 1
            assert int(b._replay.traj.len[0]) == 0                               # self.data = [] (PPO.py:133)
            if n_learn in meta["snaps"]:
                got = b.model.state_dict()
                for k in keys:    # Adam's normalised steps amplify summation-order noise of near-zero gradients (see test_learn_gpu.py)
                    d = np.abs(got[k].numpy() - z[f"snap/{n_learn}/{k}"])
                    assert d.max() <= 2e-3 and (d > 1e-4).mean() <= 0.01, (n_learn, k, float(d.max()))
    assert n_learn == meta["n_learn"]
    got = b.model.state_dict()
    for k in keys:
        d = np.abs(got[k].numpy() - z[f"final/{k}"])
        assert d.max() <= 3e-3 and (d > 1e-4).mean() <= 0.02, (k, float(d.max()), float((d > 1e-4).mean()))


def test_plugin_learn_dqn_perdqn_d3qn_behaviour():
    from reinlife_b200.Models import DQN, PERDQN, D3QN
    from brain_golden_util import golden
    obs = golden()["obs"]
    rng = np.random.default_rng(0)
    torch.manual_seed(1)

    def feed(brain, n, age0=2, **kw):
        for t in range(n):
            r = float(rng.choice([0.0, 0.2, 0.5, -3.0]))
            brain.learn(age=age0 + t, dead=False, action=int(rng.integers(8)), state=obs[rng.integers(512)], reward=r,
                        state_prime=obs[rng.integers(512)], done=False, **kw)

    # DQN.py:78-89: train() on a trigger only once the buffer holds > 1000 items: 5 optimizer steps, then target <- agent
    b = DQN(max_epi=100, train_freq=10, buffer_limit=1200)
    w0 = b.agent.state_dict()["fc1.weight"].clone()
    feed(b, 1000)
    assert int(b._dev.adam_step) == 0 and torch.equal(b.agent.state_dict()["fc1.weight"], w0)
    feed(b, 12, age0=1001)                                         # ages 1001..1012: 1010 triggers with 1010 > 1000 items
    assert int(b._dev.adam_step) == 5
    assert int(b._replay.len[0]) == 1012
    for k, v in b.agent.state_dict().items():
        assert torch.equal(v, b.target.state_dict()[k])
    a = b.get_action(obs[0], 30)
    assert 0 <= a < 8

    # PERDQN.py:188-195: train_model once the memory holds train_start (1000) items; epsilon steps per train_model
    p = PERDQN(train_freq=10, capacity=1500)
    feed(p, 999)
    assert int(p._dev.adam_step) == 0 and p.epsilon == 1.0
    feed(p, 21, age0=1001)                                         # triggers at ages 1010, 1020
    assert int(p._dev.adam_step) == 2
    assert abs(p.epsilon - (1.0 - 2 * p.epsilon_decay)) < 1e-12
    assert 0 <= p.get_action(obs[1]) < 8

    # D3QN.py:140: random.sample on a buffer shorter than a batch raises ValueError inside learn(), before any update
    d = D3QN(exploration=0, train_freq=100)
    feed(d, 3, age0=2, n_epi=1)                                    # ages 2, 3, 4: no trigger
    with pytest.raises(ValueError):
        d.learn(age=100, dead=False, action=0, state=obs[0], reward=0.0, state_prime=obs[1], done=False, n_epi=1)
    assert int(d._dev.adam_step) == 0
    feed(d, 64, age0=101, n_epi=2)                                 # 68 items, no trigger
    d.learn(age=200, dead=False, action=0, state=obs[0], reward=0.0, state_prime=obs[1], done=False, n_epi=2)
    assert int(d._dev.adam_step) == 1

"""TEST INFRASTRUCTURE.  Mints tests/golden/plugin_golden.npz: the reference's per-agent plugin call
`PPOAgent.learn(age, dead, action, state, reward, state_prime, done, prob)` (Models/PPO.py:71-77 -> put_data :113,
learn :136-162) driven call by call, the way World/entities.py:194-208 drives it: 90 calls, train_freq = 7, some of
them dead (a dead agent triggers learn() too).  Recorded: every call's arguments (observations as indices into the
`obs` pool of brain_golden.npz) and the model's weights after the last epoch of selected learn() calls.

    python oracle/make_plugin_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "plugin_golden.npz")
SNAP_LEARN = {1, 2, 5, 9}


def main():
    rh.load_reference()
    from ReinLife.Models.PPO import PPOAgent
    torch.set_num_threads(1)
    torch.manual_seed(5)
    obs = np.load(os.path.join(HERE, "..", "tests", "golden", "brain_golden.npz"))["obs"]
    rng = np.random.default_rng(42)
    agent = PPOAgent(train_freq=7)
    out = {f"w0/{k}": v.detach().numpy().copy() for k, v in agent.model.state_dict().items()}
    n_learn = [0]
    real_learn = agent.model.learn

    def learn():
        real_learn()
        n_learn[0] += 1
        if n_learn[0] in SNAP_LEARN:
            for k, v in agent.model.state_dict().items():
                out[f"snap/{n_learn[0]}/{k}"] = v.detach().numpy().copy()

    agent.model.learn = learn
    N = 90
    calls = np.zeros(N, dtype=[("age", "i4"), ("dead", "u1"), ("action", "i4"), ("s", "i4"), ("sp", "i4"), ("reward", "f8")])
    probs = np.zeros((N, 8), np.float32)
    ages = rng.integers(2, 40, 6).tolist()
    for t in range(N):
        k = t % len(ages)
        ages[k] += 1
        dead = bool(rng.random() < 0.08)
        s, sp = int(rng.integers(0, 512)), int(rng.integers(0, 512))
        with torch.no_grad():
            a, prob = agent.get_action(obs[s])
        reward = float(rng.choice([0.0, 0.2, 0.45, 0.5, -3.0, -11.0]))
        calls[t] = (ages[k], dead, a, s, sp, reward)
        probs[t] = prob.detach().numpy()
        agent.learn(age=ages[k], dead=dead, action=a, state=obs[s], reward=reward, state_prime=obs[sp], done=dead, prob=prob.detach())
        if dead:
            ages[k] = 1
    for k, v in agent.model.state_dict().items():
        out[f"final/{k}"] = v.detach().numpy().copy()
    out["calls"], out["probs"] = calls, probs
    meta = dict(train_freq=7, n_learn=n_learn[0], snaps=sorted(SNAP_LEARN), left_in_list=len(agent.model.data))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB", meta)


if __name__ == "__main__":
    main()

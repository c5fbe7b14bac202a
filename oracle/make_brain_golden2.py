"""TEST INFRASTRUCTURE.  Mints tests/golden/brain_golden2.npz from the reference's OWN DQN and PPO training code
(build container only; the reference is imported unmodified):

* train_dqn: DQNAgent.train() -> train(q, q_target, memory, optimizer) (Models/DQN.py:78-81,142-153): the five sampled
  batches of one call, and the online weights after each of the five optimizer steps;
* train_ppo/T<n>: PPO.learn() (Models/PPO.py:136-162) on a data list of n transitions (n = 1, 5, 37, 100): the
  batch, and the weights after each of the k_epoch = 3 optimizer steps.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import ref_harness as rh  # noqa: E402
from golden_util import load_cases  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "brain_golden2.npz")


def sd_np(sd, prefix):
    return {f"{prefix}/{k}": v.detach().numpy().copy() for k, v in sd.items()}


def main():
    rh.load_reference()
    from ReinLife.Models.DQN import DQNAgent
    dqn_mod = sys.modules["ReinLife.Models.DQN"]
    from ReinLife.Models.PPO import PPO
    torch.set_num_threads(1)
    rng = np.random.default_rng(11)
    obs_all = np.concatenate([c["out_obs"] for c in load_cases() if len(c["out_obs"])], 0)
    obs = obs_all[rng.choice(len(obs_all), 512, replace=False)]
    out, meta = {}, {}

    # ---------------- DQN: one train() call = 5 x (sample 32, smooth-L1, Adam)
    import random
    torch.manual_seed(321); random.seed(321)
    agent = DQNAgent(max_epi=100)
    with torch.no_grad():
        for p in agent.target.parameters():
            p.add_(0.05 * torch.randn_like(p))
    for i in range(1100):
        r = float(rng.choice([0.0, 0.2, 0.45, 0.5, -3.0, -42.0]))
        agent.memorize(obs[i % 512], int(rng.integers(8)), r, obs[(i * 7 + 3) % 512], bool(r < 0))
    out.update(sd_np(agent.agent.state_dict(), "train_dqn/w0"))
    out.update(sd_np(agent.target.state_dict(), "train_dqn/target"))
    batches, weights = [], []
    real_sample = agent.memory.sample

    def rec_sample(n):
        res = real_sample(n)
        batches.append([t.numpy().copy() for t in res])
        return res
    agent.memory.sample = rec_sample
    real_step = agent.optimizer.step

    def rec_step(*a, **k):
        r_ = real_step(*a, **k)
        weights.append({k_: v.detach().numpy().copy() for k_, v in agent.agent.state_dict().items()})
        return r_
    agent.optimizer.step = rec_step
    agent.train()
    assert len(batches) == 5 and len(weights) == 5
    for it in range(5):
        s, a, r, sp, dm = batches[it]
        p = f"train_dqn/i{it}/"
        out[p + "obs"], out[p + "action"], out[p + "reward"] = s.astype(np.float64), a[:, 0], r[:, 0].astype(np.float64)
        out[p + "next_obs"], out[p + "done_mask"] = sp.astype(np.float64), dm[:, 0].astype(np.float64)
        out.update({f"{p}w/{k}": v for k, v in weights[it].items()})
    for k, v in agent.target.state_dict().items():          # train() ends with target <- agent (DQN.py:81)
        assert torch.equal(v, agent.agent.state_dict()[k])
    meta["dqn"] = dict(lr=5e-4, gamma=dqn_mod.gamma, batch=dqn_mod.batch_size, iters=5)

    # ---------------- PPO: learn() on data lists of several lengths
    torch.manual_seed(99)
    base = PPO(153, 8, 5e-4, 0.98, 0.95, 0.1, 3)
    out.update(sd_np(base.state_dict(), "train_ppo/w0"))
    w0 = {k: v.clone() for k, v in base.state_dict().items()}
    for T in (1, 5, 37, 100):
        model = PPO(153, 8, 5e-4, 0.98, 0.95, 0.1, 3)
        model.load_state_dict(w0)
        sel = rng.integers(0, 512, T)
        s = obs[sel]; sp = obs[(sel * 5 + 1) % 512]
        with torch.no_grad():
            pi = model.pi(torch.tensor(s, dtype=torch.float), softmax_dim=1).numpy()
        a = rng.integers(0, 8, T)
        prob_a = (pi[np.arange(T), a] * rng.uniform(0.7, 1.3, T)).astype(np.float32)
        r = rng.choice([0.0, 0.2, 0.45, 0.5, -3.0, -42.0], T) / 100.0
        done = rng.random(T) < 0.15
        for t in range(T):
            model.put_data((s[t], int(a[t]), float(r[t]), sp[t], float(prob_a[t]), bool(done[t])))
        weights = []
        real_step = model.optimizer.step

        def rec_step(*a_, _m=model, _real=real_step, _w=weights, **k_):
            r_ = _real(*a_, **k_)
            _w.append({k: v.detach().numpy().copy() for k, v in _m.state_dict().items()})
            return r_
        model.optimizer.step = rec_step
        model.learn()
        assert len(weights) == 3 and model.data == []
        p = f"train_ppo/T{T}/"
        out[p + "obs"], out[p + "next_obs"], out[p + "action"] = s, sp, a.astype(np.int64)
        out[p + "reward"], out[p + "prob_a"], out[p + "done"] = r.astype(np.float64), prob_a, done
        for e in range(3):
            out.update({f"{p}e{e}/{k}": v for k, v in weights[e].items()})
    meta["ppo"] = dict(lr=5e-4, gamma=0.98, lmbda=0.95, eps_clip=0.1, k_epoch=3)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    main()

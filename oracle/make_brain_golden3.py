"""TEST INFRASTRUCTURE.  Mints tests/golden/brain_golden3.npz from the reference's OWN PERDQN code (build container only;
the reference source is imported unmodified -- only two module-level names it looks up at call time are rebound):

* `np` inside ReinLife.Models.PERDQN -> a pass-through proxy whose `array()` builds the (64, 5) object array that
  numpy 1.17 (the reference's pin, requirements.txt:2) built from the ragged mini-batch at PERDQN.py:136; numpy >= 1.24
  raises ValueError there.  Everything else is numpy 2.3 as installed.
* `random` inside the module -> a proxy that records the random.random() value behind every random.uniform(a, b)
  (CPython: a + (b - a) * random()), so the stratified sampler can be replayed.

Recorded: forward outputs with the pretrained weights; a 150-store / 3-train / 250-store (ring wrap) / 1-train run of
PERDQNAgent(capacity=300): every append_sample error and leaf, every sampled batch (tree indices, importance weights,
uniform draws), every priority update, the tree, beta and epsilon after each phase, the weights after each optimizer step.
"""
import json
import os
import random as _random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import ref_harness as rh  # noqa: E402
from golden_util import load_cases  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "brain_golden3.npz")
CAP = 300


class NpProxy:
    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def array(x, *a, **k):
        try:
            return np.array(x, *a, **k)
        except ValueError:                       # ragged list of (state, action, reward, next_state, done) tuples
            out = np.empty((len(x), len(x[0])), dtype=object)
            for i, row in enumerate(x):
                for j, v in enumerate(row):
                    out[i, j] = v
            return out


class RandomProxy:
    def __init__(self, seed):
        self._r = _random.Random(seed)
        self.us = []

    def uniform(self, a, b):
        u = self._r.random()
        self.us.append(u)
        return a + (b - a) * u

    def randrange(self, *a):
        return self._r.randrange(*a)

    def random(self):
        return self._r.random()


def sd_np(sd, prefix):
    return {f"{prefix}/{k}": v.detach().numpy().copy() for k, v in sd.items()}


def main():
    rh.load_reference()
    from ReinLife.Models.PERDQN import PERDQNAgent, DQN as RefNet
    mod = sys.modules["ReinLife.Models.PERDQN"]
    mod.np = NpProxy()
    rp = RandomProxy(2024)
    mod.random = rp
    torch.set_num_threads(1)
    rng = np.random.default_rng(5)
    obs_all = np.concatenate([c["out_obs"] for c in load_cases() if len(c["out_obs"])], 0)
    obs = obs_all[rng.choice(len(obs_all), 512, replace=False)]
    out, meta = {}, {}

    # ---------------- forward with the pretrained weights (PERDQN.py:311-323)
    net = RefNet(153, 8)
    sd = torch.load(os.path.join(rh.REF_PATH, "pretrained", "PERDQN", "PERDQN", "brain_gene_0.pt"), map_location="cpu")
    net.load_state_dict(sd)
    out.update(sd_np(sd, "fwd/w"))
    out["fwd/obs"] = obs[:64]
    with torch.no_grad():
        out["fwd/q"] = net(torch.tensor(obs[:64], dtype=torch.float)).numpy()

    # ---------------- construction: weights drawn from torch's global RNG (xavier on the online model, PERDQN.py:73-83)
    torch.manual_seed(123)
    fresh = PERDQNAgent()
    out.update(sd_np(fresh.model.state_dict(), "init/w"))
    out["init/next_rand"] = torch.rand(4).numpy()              # where the global stream stands after construction

    # ---------------- a training run
    torch.manual_seed(77)
    agent = PERDQNAgent(capacity=CAP, explore_step=50)
    agent.train_start = 100
    with torch.no_grad():
        for p in agent.target_model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    out.update(sd_np(agent.model.state_dict(), "run/w0"))
    out.update(sd_np(agent.target_model.state_dict(), "run/target"))
    mem = agent.memory
    log = {"add_err": [], "add_leaf": [], "upd_idx": [], "upd_err": []}
    real_add, real_update, real_sample = mem.add, mem.update, mem.sample

    def rec_add(error, sample):
        w = mem.tree.write
        real_add(error, sample)
        log["add_err"].append(float(error)); log["add_leaf"].append(mem.tree.tree[w + CAP - 1])
    def rec_update(idx, error):
        log["upd_idx"].append(int(idx)); log["upd_err"].append(np.float32(error))
        real_update(idx, error)
    samples = []
    def rec_sample(n):
        u0 = len(rp.us)
        res = real_sample(n)
        samples.append((list(res[1]), np.asarray(res[2], np.float64).copy(), np.asarray(rp.us[u0:])))
        return res
    mem.add, mem.update, mem.sample = rec_add, rec_update, rec_sample
    weights = []
    real_step = agent.optimizer.step
    def rec_step(*a, **k):
        r_ = real_step(*a, **k)
        weights.append({k_: v.detach().numpy().copy() for k_, v in agent.model.state_dict().items()})
        return r_
    agent.optimizer.step = rec_step

    tr = dict(state=[], action=[], reward=[], next_state=[], done=[])
    def store(n):
        for _ in range(n):
            i = len(tr["action"])
            r = float(rng.choice([0.0, 0.2, 0.45, 0.5, 0.7, -3.0, -42.0]))
            s, a, s2, d = obs[i % 512], int(rng.integers(8)), obs[(i * 7 + 3) % 512], bool(r < 0)
            agent.append_sample(s, a, r, s2, d)
            tr["state"].append(s); tr["action"].append(a); tr["reward"].append(r); tr["next_state"].append(s2); tr["done"].append(d)

    phases = []
    def snap(name):
        p = f"run/{name}/"
        out[p + "tree"] = mem.tree.tree.copy()
        out[p + "scal"] = np.array([mem.tree.write, mem.tree.n_entries, mem.beta, agent.epsilon], np.float64)
        phases.append(name)

    store(150); snap("p0_store150")
    for k in range(3):
        agent.train_model(); snap(f"p{k + 1}_train")
    store(250); snap("p4_store250")
    agent.train_model(); snap("p5_train")
    assert len(weights) == 4 and len(samples) == 4
    for k_, v in tr.items():
        out["run/" + k_] = np.asarray(v)
    out["run/add_err"] = np.asarray(log["add_err"], np.float32)
    out["run/add_leaf"] = np.asarray(log["add_leaf"], np.float64)
    out["run/upd_idx"] = np.asarray(log["upd_idx"], np.int64)
    out["run/upd_err"] = np.asarray(log["upd_err"], np.float32)
    for k, (idxs, isw, us) in enumerate(samples):
        out[f"run/sample{k}/idx"], out[f"run/sample{k}/isw"], out[f"run/sample{k}/u"] = np.asarray(idxs, np.int64), isw, us
        out.update({f"run/step{k}/{n}": v for n, v in weights[k].items()})
    meta["perdqn"] = dict(capacity=CAP, lr=1e-3, gamma=0.99, batch=64, train_start=100, eps_decay=agent.epsilon_decay,
                          eps_min=agent.epsilon_min, phases=phases, numpy=np.__version__, torch=torch.__version__)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    main()

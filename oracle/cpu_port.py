"""TEST INFRASTRUCTURE / CPU BASELINE -- never imported by reinlife_b200/.

A CPU port of the whole hot loop (act -> step -> learn -> update_env [-> top-up]) built from the oracles:
the C restatement of the World (oracle/rl_oracle.c) and the fp32 torch-CPU restatement of the PERD3QN brain
(oracle/brain_oracle.py; the reference's own network math is torch CPU too).  It follows the same N-world
semantics as the CUDA path (per-world replay rings, every train() trigger is a 64-row event, per-event gradients
averaged into one Adam step per brain per step), so bench.py can time "the reference's algorithm on host cores"
on exactly the GPU arm's workload:  bench.py's `cpu_baseline` leg and `--impl reference` arm (kind = "port").

It is far faster than the unmodified Python reference (the World step is compiled C here, np.vectorize there);
BASELINE.md section 2 has the survey-time numbers of the real reference for context.
"""
import time

import numpy as np
import torch

from . import brain_oracle as bo
from .world_oracle import OracleWorlds

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    z = z.astype(np.uint64)
    z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def world_keys(seed, ids):
    with np.errstate(over="ignore"):
        return _mix64(np.uint64(seed) ^ _mix64(np.asarray(ids, np.uint64) + np.uint64(0x9E3779B97F4A7C15)))


def draws(keys, step, site, idx):
    """Vectorised include/rl_rng.h: keys uint64[n], idx uint32-like[n] -> uint64[n]."""
    with np.errstate(over="ignore"):
        x = _mix64(keys + np.uint64(step) * np.uint64(0xD1342543DE82EF95) + np.uint64(0x9E3779B97F4A7C15))
        return _mix64(x ^ ((np.uint64(site) << np.uint64(32)) | np.asarray(idx, np.uint64)))


def uniform(bits):
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def below(bits, n):
    return ((bits >> np.uint64(32)) * np.uint64(n)) >> np.uint64(32)


class CpuPort:
    def __init__(self, n_worlds, n_brains=2, height=30, width=30, max_agents=100, seed=0, world_id0=0, capacity=10000,
                 exploration=0, train_freq=20, soft_update_freq=200, lr=1e-3, gamma=0.99, saturate_to=100, training=True):
        from reinlife_b200.Models import packing   # only the nn.Linear init helper (host-side, no CUDA)
        self.nw, self.G, self.saturate_to, self.training = n_worlds, n_brains, saturate_to, training
        self.ow = OracleWorlds(n_worlds, height, width, n_brains, max_agents, seed=seed, world_id0=world_id0)
        self.keys = world_keys(seed, np.arange(world_id0, world_id0 + n_worlds))
        self.seed, self.cap = seed, capacity
        self.exploration, self.train_freq, self.soft_update_freq, self.lr, self.gamma = exploration, train_freq, soft_update_freq, lr, gamma
        self.sd = [{k: v.numpy().copy() for k, v in packing.default_init(packing.DUELING).items()} for _ in range(n_brains)]
        self.tgt = [{k: v.copy() for k, v in sd.items()} for sd in self.sd]
        self.m = [{k: np.zeros_like(v) for k, v in sd.items()} for sd in self.sd]
        self.v = [{k: np.zeros_like(v) for k, v in sd.items()} for sd in self.sd]
        self.adam_t = [0] * n_brains
        self.eps = [0.9 if training else 0.0] * n_brains
        self.seen = [0] * n_brains
        if training:
            S = capacity
            self.rp = [[dict(obs=np.zeros((S, 153), np.float32), nobs=np.zeros((S, 153), np.float32), act=np.zeros(S, np.int64),
                             rew=np.zeros(S, np.float32), done=np.zeros(S, np.float32), prio=np.zeros(S, np.float32), len=0, pos=0)
                        for _ in range(n_brains)] for _ in range(n_worlds)]
        self.ow.reset()
        if saturate_to:
            self.ow.top_up(saturate_to)
        self.state = self.ow.obs.copy()

    def step_loop(self, n_epi):
        ow, nw, G = self.ow, self.nw, self.G
        n = ow.n.copy()
        agent_steps = int(n.sum())
        t_act = ow.t + 1
        # ---- act: one batched forward per brain (reference: B=1 per agent, PERD3QN.py:81-89, 204-210)
        gene = ow.rec["gene"]
        for g in range(G):
            ws, ss = np.nonzero((np.arange(ow.S)[None, :] < n[:, None]) & (gene == g))
            if len(ws) == 0:
                continue
            if self.training and n_epi > self.seen[g]:
                if self.eps[g] > 0.05:
                    self.eps[g] *= 0.99
                self.seen[g] = n_epi
            q = bo.dueling_forward(self.sd[g], self.state[ws, ss].astype(np.float32), per_row_mean=True)
            a = q.argmax(1)
            u = uniform(draws(self.keys[ws], t_act, 20, ss))
            rnd = below(draws(self.keys[ws], t_act, 21, ss), 8).astype(np.int64)
            a = np.where(u > self.eps[g], a, rnd)
            ow.rec["action"][ws, ss] = a.astype(np.int8)
        # ---- step
        ow.step()
        # ---- learn
        if self.training:
            self._learn(n_epi)
        # ---- update_env (+ saturated top-up)
        ow.update()
        if self.saturate_to:
            ow.top_up(self.saturate_to)
        self.state = ow.obs.copy()
        return agent_steps

    def _learn(self, n_epi):
        ow, G = self.ow, self.G
        prime = ow.obs
        events = [[] for _ in range(G)]
        ev_meta = [[] for _ in range(G)]
        for w in range(self.nw):
            n = int(ow.n[w])
            rec = ow.rec[w, :n]
            for g in range(G):
                sel = np.nonzero((rec["gene"] == g) & (rec["age"] > 1))[0]
                if len(sel) == 0:
                    continue
                rp = self.rp[w][g]
                maxp = rp["prio"].max() if rp["len"] > 0 else 1.0                      # PERD3QN.py:147
                pos = (rp["pos"] + np.arange(len(sel))) % self.cap
                rp["obs"][pos] = self.state[w, rec["prev_slot"][sel]]
                rp["nobs"][pos] = prime[w, sel]
                rp["act"][pos] = rec["action"][sel]
                rp["rew"][pos] = ow.reward[w, sel]
                rp["done"][pos] = (rec["flags"][sel] & 32) != 0
                rp["prio"][pos] = maxp
                rp["pos"] = int((rp["pos"] + len(sel)) % self.cap)
                rp["len"] = int(min(self.cap, rp["len"] + len(sel)))
                if n_epi > self.exploration:
                    trig = sel[(rec["age"][sel] % self.train_freq == 0) | ((rec["flags"][sel] & 32) != 0)]
                    if len(trig):
                        wts = bo.per_weight(rp["prio"][:rp["len"]]).astype(np.float64)
                        cdf = np.cumsum(wts)
                        for k in range(len(trig)):
                            u = uniform(draws(np.full(64, self.keys[w]), ow.t, 30, k * 64 + np.arange(64)))
                            idx = np.minimum(np.searchsorted(cdf, u * cdf[-1], side="right"), rp["len"] - 1)
                            events[g].append((rp["obs"][idx], rp["act"][idx], rp["rew"][idx], rp["nobs"][idx], rp["done"][idx]))
                            ev_meta[g].append((w, idx))
        for g in range(G):
            if events[g]:
                grads, _, prios = bo.dueling_batched_update(self.sd[g], self.tgt[g], events[g], self.gamma)
                self.adam_t[g] += 1
                bo.adam_step(self.sd[g], grads, self.m[g], self.v[g], self.adam_t[g], self.lr)
                for (w, idx), pr in zip(ev_meta[g], prios):
                    self.rp[w][g]["prio"][idx] = pr
            if n_epi > self.exploration and n_epi % self.soft_update_freq == 0:
                self.tgt[g] = {k: v.copy() for k, v in self.sd[g].items()}


def run_sample(n_worlds, steps, warmup, seed=0, world_id0=0, capacity=10000, training=True, saturate_to=100, **kw):
    """-> (agent_steps, seconds) over `steps` timed iterations of the hot loop on `n_worlds` worlds, one thread."""
    torch.set_num_threads(1)
    port = CpuPort(n_worlds, seed=seed, world_id0=world_id0, capacity=capacity, training=training,
                   saturate_to=saturate_to, **kw)
    for i in range(warmup):
        port.step_loop(i + 1)
    t0 = time.perf_counter()
    total = 0
    for i in range(steps):
        total += port.step_loop(warmup + i + 1)
    return total, time.perf_counter() - t0


def run_parallel(procs, n_worlds_each, steps, warmup, seed=0, timeout=1200, **kw):
    """`procs` independent single-thread worker processes (python -m oracle.cpu_port ...), disjoint world ranges.
    -> (agent_steps_total, max_seconds)"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    ps = []
    for p in range(procs):
        spec = dict(n_worlds=n_worlds_each, steps=steps, warmup=warmup, seed=seed, world_id0=p * n_worlds_each, **kw)
        ps.append(subprocess.Popen([sys.executable, "-m", "oracle.cpu_port", json.dumps(spec)], cwd=root, env=env,
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    tot, mx = 0, 0.0
    for p in ps:
        out, err = p.communicate(timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("cpu_port worker failed: " + err[-2000:])
        r = json.loads(out.strip().splitlines()[-1])
        tot += r["agent_steps"]
        mx = max(mx, r["seconds"])
    return tot, mx


if __name__ == "__main__":
    import json
    import sys
    spec = json.loads(sys.argv[1])
    a, sec = run_sample(**spec)
    print(json.dumps(dict(agent_steps=a, seconds=sec)))

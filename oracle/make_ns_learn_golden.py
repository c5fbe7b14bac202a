"""TEST INFRASTRUCTURE.  Mints tests/golden/ns_learn_golden.npz: ONE world of the UNMODIFIED reference with
static_families=False and learning ON -- lineage brains shared by offspring (World/environment.py:506-507), deep-copied
from a random best agent at every _produce (:541-547, incl. replay memory and Adam state) and mutated (entities.py:210-213),
`best_agents` maintained on Agent.fitness (:728-739) -- driven through the reference's loop body (Helpers/trainer.py:85-99)
with teacher-forced actions.  Recorded: initial weights, the actions, for every train() event the gene of the agent that
triggered it + the 64 ring positions np.random.choice drew + the loss, every _produce event, and at the end the eval-net
weights / Adam step count / ring fill of every brain still referenced (live lineages and best-table entries).

    python oracle/make_ns_learn_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "ns_learn_golden.npz")
CFG = dict(height=9, width=9, n_genes=2, max_agents=12, seed=41, world=2, steps=260,
           exploration=3, train_freq=3, capacity=120, soft_update_freq=5, lr=1e-3, gamma=0.99)


def main():
    rh.load_reference()
    from ReinLife.Models.PERD3QN import PERD3QNAgent, PrioritizedReplayBuffer
    torch.set_num_threads(1)
    torch.manual_seed(5); np.random.seed(5)
    c = CFG
    brains = [PERD3QNAgent(exploration=c["exploration"], train_freq=c["train_freq"], capacity=c["capacity"],
                           soft_update_freq=c["soft_update_freq"], learning_rate=c["lr"], gamma=c["gamma"]) for _ in range(2)]
    out = {}
    for g, b in enumerate(brains):
        for k, v in b.eval_net.state_dict().items():
            out[f"w0/{g}/{k}"] = v.detach().numpy().copy()
    ev_idx, ev_loss = [], []
    real_sample = PrioritizedReplayBuffer.sample          # class-level hooks survive copy.deepcopy of a brain

    def sample(self, bs):
        res = real_sample(self, bs)
        ev_idx.append(np.asarray(res[5], np.int32).copy())
        return res
    PrioritizedReplayBuffer.sample = sample
    real_train = PERD3QNAgent.train

    def train(self):
        real_loss = self.loss_fn

        def loss_fn(a, t):
            v = real_loss(a, t)
            ev_loss.append(float(v))
            return v
        self.loss_fn = loss_fn
        try:
            real_train(self)
        finally:
            self.loss_fn = real_loss
    PERD3QNAgent.train = train

    w = rh.RefWorld(brains, seed=c["seed"], world=c["world"], width=c["width"], height=c["height"], max_agents=c["max_agents"],
                    static_families=False, training=False)
    rng = np.random.default_rng(7)
    w.reset()
    actions, counts, ev_gene, ev_step, produced = [], [], [], [], []
    for n_epi in range(c["steps"] + 1):
        n = len(w.env.agents)
        a = rng.integers(0, 8, size=n)
        actions.append(a.astype(np.int8)); counts.append(n)
        w.force_actions(a)
        w.step()
        for agent in w.env.agents:                        # Helpers/trainer.py:95-96
            before = len(ev_idx)
            agent.learn(n_epi=n_epi)
            if len(ev_idx) > before:
                ev_gene.append(agent.gene); ev_step.append(n_epi)
        mg = w.env.max_gene
        w.update_env(n_epi)
        if w.env.max_gene > mg:
            produced.append((n_epi, w.env.max_gene, rh.CTX.last_choice))
    env = w.env
    refs = {}
    for agent in env.agents:
        refs[int(agent.gene)] = agent.brain
    for b in env.best_agents:
        if getattr(b, "_serial", 0) >= 0:
            refs.setdefault(int(b.gene), b.brain)
    final = []
    for g, br in sorted(refs.items()):
        for k, v in br.eval_net.state_dict().items():
            out[f"final/{g}/{k}"] = v.detach().numpy().copy()
        for k, v in br.target_net.state_dict().items():
            out[f"final_target/{g}/{k}"] = v.detach().numpy().copy()
        steps = 0
        st = br.optimizer.state
        if len(st):
            steps = int(next(iter(st.values()))["step"])
        final.append((g, steps, len(br.buffer.memory), int(br.buffer.pos)))
    out["actions"] = np.concatenate(actions); out["counts"] = np.array(counts, np.int32)
    out["ev_idx"] = np.stack(ev_idx) if ev_idx else np.zeros((0, 64), np.int32)
    out["ev_loss"] = np.array(ev_loss); out["ev_gene"] = np.array(ev_gene, np.int32); out["ev_step"] = np.array(ev_step, np.int32)
    out["produced"] = np.array(produced, np.int32).reshape(-1, 3)
    meta = dict(c, final=final, max_gene=int(env.max_gene), n_events=len(ev_idx))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB", {k: meta[k] for k in ("final", "max_gene", "n_events")}, "produced", len(produced))


if __name__ == "__main__":
    main()

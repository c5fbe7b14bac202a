"""TEST INFRASTRUCTURE.  Mints tests/golden/seq_golden.npz: ONE world of the UNMODIFIED reference run through the
reference's own trainer loop body (Helpers/trainer.py:85-99) with learning ON -- `for agent in env.agents:
agent.learn(n_epi=n_epi)` -> World/entities.py:194-208 -> PERD3QNAgent.learn / memorize / train
(Models/PERD3QN.py:91-125) -- so that the per-agent order of effects (store -> sample -> train -> priorities -> Adam ->
target sync, each train() seeing the weights and the ring the previous agent's train() left) is pinned.

Teacher-forced: the actions (uniform random, like the world goldens) and -- recorded, not forced -- the ring positions
np.random.choice drew inside every buffer.sample (the samplers have their own bit-exact tests; here they are replayed).
Recorded: initial weights of both brains, per step the forced actions, per train() event (brain, 64 indices, loss,
64 new priorities), the eval-net weights after selected optimizer steps, and both rings' priorities at the end.

    python oracle/make_seq_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "seq_golden.npz")
CFG = dict(height=10, width=10, n_genes=2, max_agents=20, seed=31, world=4, steps=200, top_up=12,
           exploration=3, train_freq=4, capacity=150, soft_update_freq=7, lr=1e-3, gamma=0.99)
SNAP_AT = {1, 2, 3, 10, 50, 100, 200, 300, 400}


def main():
    rh.load_reference()
    from ReinLife.Models.PERD3QN import PERD3QNAgent
    torch.set_num_threads(1)
    torch.manual_seed(77)
    np.random.seed(77)
    c = CFG
    brains = [PERD3QNAgent(exploration=c["exploration"], train_freq=c["train_freq"], capacity=c["capacity"],
                           soft_update_freq=c["soft_update_freq"], learning_rate=c["lr"], gamma=c["gamma"]) for _ in range(2)]
    out = {}
    for g, b in enumerate(brains):
        for k, v in b.eval_net.state_dict().items():
            out[f"w0/{g}/{k}"] = v.detach().numpy().copy()
    events = []          # (step, brain, adam_step_of_that_brain)
    ev_idx, ev_loss, ev_prio = [], [], []
    adam_steps = [0, 0]
    snaps = []

    for g, b in enumerate(brains):
        real_sample, real_update, real_loss = b.buffer.sample, b.buffer.update_priorities, b.loss_fn
        real_step = b.optimizer.step

        def sample(bs, _r=real_sample, _g=g):
            res = _r(bs)
            ev_idx.append(np.asarray(res[5], np.int32).copy())
            return res

        def update(indices, priorities, _r=real_update):
            ev_prio.append(np.asarray(priorities, np.float32).copy())
            return _r(indices, priorities)

        def loss_fn(a, t, _r=real_loss):
            val = _r(a, t)
            ev_loss.append(float(val))
            return val

        def step(*a, _r=real_step, _g=g, _b=b, **kw):
            res = _r(*a, **kw)
            adam_steps[_g] += 1
            events.append((cur["n_epi"], _g, adam_steps[_g]))
            if adam_steps[_g] in SNAP_AT:
                snaps.append((_g, adam_steps[_g]))
                for k, v in _b.eval_net.state_dict().items():
                    out[f"snap/{_g}/{adam_steps[_g]}/{k}"] = v.detach().numpy().copy()
            return res

        b.buffer.sample, b.buffer.update_priorities, b.loss_fn, b.optimizer.step = sample, update, loss_fn, step

    cur = {"n_epi": 0}
    w = rh.RefWorld(brains, seed=c["seed"], world=c["world"], width=c["width"], height=c["height"],
                    max_agents=c["max_agents"], training=False)
    rng = np.random.default_rng(99)
    w.reset()
    w.top_up(c["top_up"])
    actions, counts = [], []
    for n_epi in range(c["steps"] + 1):
        cur["n_epi"] = n_epi
        n = len(w.env.agents)
        a = rng.integers(0, 8, size=n)
        actions.append(a.astype(np.int8)); counts.append(n)
        w.force_actions(a)
        w.step()
        for agent in w.env.agents:                       # Helpers/trainer.py:95-96
            agent.learn(n_epi=n_epi)
        w.update_env(n_epi)
        w.top_up(c["top_up"])
    for g, b in enumerate(brains):
        for k, v in b.eval_net.state_dict().items():
            out[f"final/{g}/{k}"] = v.detach().numpy().copy()
        for k, v in b.target_net.state_dict().items():
            out[f"final_target/{g}/{k}"] = v.detach().numpy().copy()
        out[f"final_prio/{g}"] = b.buffer.priorities.copy()
        out[f"final_pos/{g}"] = np.array([b.buffer.pos, len(b.buffer.memory)], np.int32)
    out["actions"] = np.concatenate(actions)
    out["counts"] = np.array(counts, np.int32)
    out["events"] = np.array(events, np.int32)
    out["ev_idx"] = np.stack(ev_idx)
    out["ev_loss"] = np.array(ev_loss, np.float64)
    out["ev_prio"] = np.stack(ev_prio)
    meta = dict(c, adam_steps=adam_steps, snaps=snaps)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB; adam steps", adam_steps, "events", len(events), "snaps", snaps)


if __name__ == "__main__":
    main()

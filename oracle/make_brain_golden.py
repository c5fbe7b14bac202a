"""TEST INFRASTRUCTURE.  Mints tests/golden/brain_golden.npz from the reference's OWN network modules and
pretrained weights (build container only):

* forward goldens: pretrained PERD3QN / D3QN / DQN / PPO state_dicts (pretrained/**/brain_gene_*.pt) applied
  by the reference classes (DuelingDDQN.forward, dueling_ddqn.forward, Qnet.forward, PPO.pi / PPO.v) to real
  observations taken from tests/golden/world_golden.npz;
* train goldens: three consecutive PERD3QNAgent.train() / D3QNAgent.train() calls (Models/PERD3QN.py:94-115,
  Models/D3QN.py:97-116) on a buffer filled with those observations: the sampled batch, loss inputs,
  priorities and the eval-net weights after every Adam step are recorded.
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

OUT = os.path.join(HERE, "..", "tests", "golden", "brain_golden.npz")
PRE = os.path.join(rh.REF_PATH, "pretrained")


def sd_np(sd, prefix):
    return {f"{prefix}/{k}": v.detach().numpy().copy() for k, v in sd.items()}


def main():
    rh.load_reference()
    from ReinLife.Models.PERD3QN import PERD3QNAgent, DuelingDDQN
    from ReinLife.Models.D3QN import D3QNAgent, dueling_ddqn
    from ReinLife.Models.DQN import Qnet
    from ReinLife.Models.PPO import PPO
    torch.set_num_threads(1)
    rng = np.random.default_rng(7)
    obs_all = np.concatenate([c["out_obs"] for c in load_cases() if len(c["out_obs"])], 0)
    pick = rng.choice(len(obs_all), 512, replace=False)
    obs = obs_all[pick]                      # float64 [512,153] real observations
    out = {"obs": obs}
    meta = {}

    # ---- forwards with pretrained weights
    net = DuelingDDQN(153, 8)
    net.load_state_dict(torch.load(os.path.join(PRE, "PERD3QN", "Static Families", "PERD3QN", "brain_gene_1.pt")))
    out.update(sd_np(net.state_dict(), "perd3qn"))
    with torch.no_grad():
        out["perd3qn_q_rows"] = np.concatenate([net.forward(torch.FloatTensor(np.expand_dims(o, 0))).numpy() for o in obs], 0)
        out["perd3qn_q_batch64"] = net.forward(torch.FloatTensor(obs[:64])).numpy()
    net = dueling_ddqn(153, 8)
    net.load_state_dict(torch.load(os.path.join(PRE, "D3QN", "D3QN", "brain_gene_0.pt")))
    out.update(sd_np(net.state_dict(), "d3qn"))
    with torch.no_grad():
        out["d3qn_q_rows"] = np.concatenate([net.forward(torch.FloatTensor(np.expand_dims(o, 0))).numpy() for o in obs], 0)
    net = Qnet(153)
    net.load_state_dict(torch.load(os.path.join(PRE, "DQN", "DQN", "brain_gene_0.pt")))
    out.update(sd_np(net.state_dict(), "dqn"))
    with torch.no_grad():
        out["dqn_q_rows"] = np.stack([net.forward(torch.from_numpy(o).float()).numpy() for o in obs], 0)
    net = PPO(153, 8, 5e-4, 0.98, 0.95, 0.1, 3)
    net.load_state_dict(torch.load(os.path.join(PRE, "PPO", "PPO", "brain_gene_0.pt")))
    out.update(sd_np(net.state_dict(), "ppo"))
    with torch.no_grad():
        out["ppo_pi_rows"] = np.stack([net.pi(torch.from_numpy(o).float()).numpy() for o in obs], 0)
        out["ppo_v_rows"] = np.stack([net.v(torch.from_numpy(o).float()).numpy() for o in obs], 0)

    # ---- train() goldens
    for name, cls in (("perd3qn", PERD3QNAgent), ("d3qn", D3QNAgent)):
        torch.manual_seed(123)
        np.random.seed(123)
        import random
        random.seed(123)
        agent = cls(exploration=0, gamma=0.99)
        # make the target differ from eval so the TD target is not degenerate
        with torch.no_grad():
            for p in agent.target_net.parameters():
                p.add_(0.05 * torch.randn_like(p))
        n_tr = 300
        for i in range(n_tr):
            r = float(rng.choice([0.0, 0.2, 0.45, 0.5, -3.0, -42.0]))
            d = bool(r < 0)
            agent.memorize(obs[i], int(rng.integers(8)), r, obs[(i * 7 + 3) % 512], d)
        out.update(sd_np(agent.eval_net.state_dict(), f"train_{name}/w0"))
        out.update(sd_np(agent.target_net.state_dict(), f"train_{name}/target"))
        recorded = []
        real_sample = agent.buffer.sample

        def rec_sample(bs, _real=real_sample, _rec=recorded):
            res = _real(bs)
            _rec.append(res)
            return res
        agent.buffer.sample = rec_sample
        for step in range(3):
            prio_before = agent.buffer.priorities.copy() if name == "perd3qn" else None
            agent.train()
            res = recorded[-1]
            out[f"train_{name}/s{step}/obs"] = np.asarray(res[0], np.float64)
            out[f"train_{name}/s{step}/action"] = np.asarray(res[1], np.int64)
            out[f"train_{name}/s{step}/reward"] = np.asarray(res[2], np.float64)
            out[f"train_{name}/s{step}/next_obs"] = np.asarray(res[3], np.float64)
            out[f"train_{name}/s{step}/done"] = np.asarray(res[4], np.float64)
            if name == "perd3qn":
                idx = np.asarray(res[5])
                out[f"train_{name}/s{step}/indices"] = idx
                out[f"train_{name}/s{step}/prio_before"] = prio_before
                out[f"train_{name}/s{step}/prio_after"] = agent.buffer.priorities.copy()
            out.update(sd_np(agent.eval_net.state_dict(), f"train_{name}/s{step}/w"))
        meta[name] = dict(lr=1e-3, gamma=0.99, batch=64, n_transitions=n_tr)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    main()

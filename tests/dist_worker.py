"""world_size-2 gloo worker (launched by tests/test_dist_cpu.py through torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    dist.init_process_group("gloo")
    rank, ws = dist.get_rank(), dist.get_world_size()
    from reinlife_b200.sharding import shard_worlds, allreduce_grads
    from reinlife_b200.Models import packing
    from oracle import brain_oracle as bo
    from brain_golden_util import golden, state_dict
    # 1. shard plan: contiguous, disjoint, complete
    n_local, w0 = shard_worlds(12, rank, ws)
    ids = torch.arange(w0, w0 + n_local)
    allids = [torch.zeros_like(ids) for _ in range(ws)]
    dist.all_gather(allids, ids)
    assert torch.cat(allids).tolist() == list(range(12))
    try:
        shard_worlds(7, rank, ws)
        raise SystemExit("expected ValueError")
    except ValueError:
        pass
    # 2. sharded [grad | count] all-reduce + Adam == single-process update on all events
    z = golden()
    rng = np.random.default_rng(5)
    w0sd, tgt = state_dict("train_perd3qn/w0"), state_dict("train_perd3qn/target")
    obs = z["obs"]
    events = []
    for e in range(6):
        i = rng.integers(0, 512, 64); j = rng.integers(0, 512, 64)
        r = rng.choice([0.0, 0.2, 0.5, -3.0], 64)
        events.append((obs[i], rng.integers(0, 8, 64), r, obs[j], (r < 0).astype(np.float64)))
    mine = events[rank::ws]
    keys = list(w0sd.keys())
    sizes = [w0sd[k].size for k in keys]

    def flat_sum(evs):
        acc = np.zeros(sum(sizes) + 1, np.float64)
        for ev in evs:
            g, _, _ = bo.dueling_event_grads(w0sd, tgt, *ev, 0.99)
            acc[:-1] += np.concatenate([g[k].reshape(-1) for k in keys])
            acc[-1] += 1
        return acc
    buf_a = torch.from_numpy(flat_sum(mine))
    buf_b = torch.from_numpy(flat_sum(mine)) * 2          # a second "brain" to exercise the concatenated path
    allreduce_grads([buf_a, buf_b])
    want = flat_sum(events)
    assert buf_a[-1].item() == 6 and buf_b[-1].item() == 12
    np.testing.assert_allclose(buf_a.numpy(), want, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(buf_b.numpy(), 2 * want, rtol=1e-12, atol=1e-12)
    # identical Adam step on every rank
    def adam_from(buf):
        w = {k: v.copy() for k, v in w0sd.items()}
        m = {k: np.zeros_like(v) for k, v in w.items()}; v2 = {k: np.zeros_like(v) for k, v in w.items()}
        g, o = {}, 0
        for k, n in zip(keys, sizes):
            g[k] = (buf[o:o + n] / buf[-1]).reshape(w[k].shape).astype(np.float32); o += n
        bo.adam_step(w, g, m, v2, 1, 1e-3)
        return np.concatenate([w[k].reshape(-1) for k in keys])
    mine_w = torch.from_numpy(adam_from(buf_a.numpy()))
    ws_w = [torch.zeros_like(mine_w) for _ in range(ws)]
    dist.all_gather(ws_w, mine_w)
    assert all(torch.equal(ws_w[0], x) for x in ws_w)
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()

"""2-GPU check: sharded learn step (NCCL all-reduce of [grad|count]) == single-GPU learn step on all worlds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist

def run(n_worlds, dist_on):
    import reinlife_b200 as rl
    from reinlife_b200.Models import PERD3QN
    torch.manual_seed(0)
    brains = [PERD3QN(exploration=0, train_freq=3, capacity=128), PERD3QN(exploration=0, train_freq=3, capacity=128)]
    env = rl.Environment(width=12, height=12, brains=brains, max_agents=40, print_results=False, training=True,
                         n_worlds=n_worlds, seed=4)
    env.reset(); env.top_up(40)
    for n_epi in range(1, 9):
        env.act(n_epi); env.step(); env.learn(n_epi); env.update_env(n_epi); env.top_up(40)
    torch.cuda.synchronize()
    return [b._dev.params.clone() for b in brains], env

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
params_sharded, env = run(16, True)
# every rank must hold identical weights
for p in params_sharded:
    ref = p.clone(); dist.broadcast(ref, 0)
    assert torch.equal(ref, p), "ranks diverged"
dist.barrier()
if dist.get_rank() == 0:
    torch.save([p.cpu() for p in params_sharded], "/tmp/sharded.pt")
dist.destroy_process_group()
if local == 0:
    # same 16 worlds on one GPU, no process group
    params_single, _ = run(16, False)
    sh = torch.load("/tmp/sharded.pt")
    for a, b in zip(sh, params_single):
        d = (a - b.cpu()).abs().max().item()
        print("max |sharded - single| =", d)
        assert d < 5e-5
    print("2-GPU sharded learn == single-GPU learn (fp summation order only)")

#!/usr/bin/env python
"""bench.py -- ReinLife hot-path throughput on B200: agent*steps/s for saturated 30x30x100-agent worlds with
PERD3QN x2 training (BASELINE.json configs[2] per GPU; worlds are sharded across GPUs, weak scaling).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        bench.py --gpus 8 --steps 20 --warmup 5
    python bench.py --impl reference --gpus 1 --steps 20 --warmup 5    # the unmodified Python reference (baseline/_ref), all host cores

One step = one pass of the reference's trainer loop body (Helpers/trainer.py:85-99) over every world:
act (batched get_action) -> env.step -> learn (store, PER sample, one 64-row train() event per trigger, Adam)
-> env.update_env -> saturated top-up.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 30
TARGET = 100
METRIC = "agent_steps_per_sec"
UNIT = "agent*step/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--worlds-per-gpu", type=int, default=4096)
    ap.add_argument("--capacity", type=int, default=10000)
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c5"],
                    help="c3 = BASELINE.json configs[2]/[3] (the metric's configuration, default); c2 = configs[1] (256 worlds, DQN x1 "
                         "inference, tester loop); c5 = configs[4] brain mix and grid with static_families=False at the scale the "
                         "exact per-lineage brain pools serve (secondary lines, single GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the additional fixed-4096-worlds (strong scaling) measurement")
    ap.add_argument("--precision", default="fp16", choices=["tf32", "fp16", "fp32"],
                    help="train() events: tcgen05 with fp16 operands (default) or tf32 operands (both 11 significant bits, fp32 accumulate), or fp32 CUDA-core FMA")
    ap.add_argument("--cpu-worlds-per-core", type=int, default=2)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--ref-steps", type=int, default=30, help="timed steps per process of the unmodified reference (cpu_baseline)")
    return ap.parse_args()


def workload_name(args):
    return (f"{args.worlds_per_gpu} worlds/GPU, 30x30, saturated to 100 agents/world, PERD3QNx2 training "
            f"(exploration=0, train_freq=20, batch 64, capacity {args.capacity}), static_families=True")


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_arm(args, steps, warmup, procs=None):
    from oracle import cpu_port
    if not procs:
        try:
            procs = len(os.sched_getaffinity(0))          # the cores this process may actually use (cgroup / taskset aware)
        except AttributeError:
            procs = os.cpu_count() or 1
        procs = max(1, min(procs, 128))                   # bounded: one python + torch worker process per core
    a, sec = cpu_port.run_parallel(procs, args.cpu_worlds_per_core, steps, warmup, seed=0, capacity=args.capacity,
                                   exploration=0, train_freq=20, saturate_to=TARGET)
    sample = (f"{procs} single-thread processes x {args.cpu_worlds_per_core} worlds x {steps} steps "
              f"(+{warmup} warm-up) of the same loop: C world oracle + torch-CPU fp32 PERD3QN oracle")
    return a / sec, procs, sample, sec


def ref_python_arm(steps, warmup, procs=0):
    """The UNMODIFIED Python reference (baseline/_ref, baseline/run_ref.py; nothing of reinlife_b200 / oracle on that
    path): one process per usable host core, each one saturated 30x30x100-agent world with PERD3QNx2 training, the
    reference's own loop body.  The reference is sequential per world, so its per-core rate does not depend on how many
    worlds a core is given; the aggregate over all cores is the whole box's CPU throughput for this workload."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import run_ref
    if not run_ref.available():
        return None
    t0 = time.perf_counter()
    value, procs, res = run_ref.run_parallel(procs, steps, warmup)
    sample = (f"{procs} single-thread processes x 1 world x {steps} steps (+{warmup} warm-up) of the unmodified reference "
              f"(baseline/_ref): get_action -> env.step -> learn -> update_env per agent, PERD3QNx2 exploration=0, top-up excluded")
    return value, procs, sample, max(r["sec"] for r in res), time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = "reference"
    ref = ref_python_arm(max(1, args.steps), max(0, args.warmup))
    if ref is not None:
        val, procs, sample, sec, _ = ref
    else:                                   # baseline/_ref did not travel: the oracle port of the same loop
        kind = "port"
        val, procs, sample, sec = cpu_arm(args, max(1, args.steps), max(0, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "grid": [H, W], "agents_per_world": TARGET},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- post-timing self-check
def parity_check(env, brains, precision):
    """Re-run the LAST step's train() events of every brain (same EVENT list, same sampled ring positions, current
    weights) through the benchmarked event kernel and through the fp32 CUDA-core kernel (the path held to the
    reference at atol 1e-5 in tests/test_learn_gpu.py) and report the largest disagreement: summed gradients per
    tensor relative to that tensor's gradient scale, per-event loss and priorities relative.  Outside the timed
    regions; the tolerance is the one of tests/test_scale_gpu.py."""
    import ctypes as C
    import torch
    from reinlife_b200 import _lib
    w, lib = env.world, env.world.lib
    st = w._stream()
    out = {"kernel_vs": "k_learn_dueling (fp32 FMA)", "tol_grad": 1e-2, "tol_loss": 2e-2, "events": 0,
           "max_rel_grad_err": 0.0, "max_rel_loss_err": 0.0, "max_rel_prio_err": 0.0}
    if precision == "fp32":
        out.update(ok=True, note="benchmarked kernel is the fp32 kernel itself")
        return out
    with torch.cuda.device(env.device):
        for g, b in enumerate(brains):
            dev = b._dev
            d = dev.dims
            n_ev = int(env.rows.total[g * _lib.N_ROW_KINDS + _lib.ROWS_EVENT])
            if n_ev == 0:
                continue
            res = {}
            for mode in (precision, "fp32"):
                if mode == "fp32":
                    _lib.check(lib.rl_brain_learn(C.byref(w.cfg), C.byref(env.rows.bufs), C.c_int32(g), C.byref(b._replay.bufs),
                                                  C.c_void_p(dev.sample_idx.data_ptr()), C.byref(dev.learn_bufs), st))
                elif mode == "fp16":
                    _lib.check(lib.rl_brain_learn_p(C.byref(w.cfg), C.byref(env.rows.bufs), C.c_int32(g), C.byref(b._replay.bufs),
                                                    C.c_void_p(dev.sample_idx.data_ptr()), C.byref(dev.learn_bufs),
                                                    C.c_void_p(dev.wimg_eh.data_ptr()), C.c_void_p(dev.wimg_th.data_ptr()), st))
                else:
                    _lib.check(lib.rl_brain_learn_tc(C.byref(w.cfg), C.byref(env.rows.bufs), C.c_int32(g), C.byref(b._replay.bufs),
                                                     C.c_void_p(dev.sample_idx.data_ptr()), C.byref(dev.learn_bufs),
                                                     C.c_void_p(dev.wimg_e.data_ptr()), C.c_void_p(dev.wimg_t.data_ptr()), st))
                torch.cuda.synchronize()
                res[mode] = (dev.grad[:d.n_train].double() * dev.mask.double(), dev.loss[:n_ev].double().clone(),
                             dev.new_prio[:n_ev].double().clone(), float(dev.grad[d.n_train]))
            a, ref = res[precision], res["fp32"]
            assert a[3] == ref[3] == n_ev, (a[3], ref[3], n_ev)
            for lo, hi in ((0, d.off_b1), (d.off_b1, d.off_w2t), (d.off_w2t, d.off_b2), (d.off_b2, d.off_wh),
                           (d.off_wh, d.off_bh), (d.off_bh, d.off_bh + 9)):
                scale = float(ref[0][lo:hi].abs().max())
                out["max_rel_grad_err"] = max(out["max_rel_grad_err"], float((a[0][lo:hi] - ref[0][lo:hi]).abs().max()) / max(scale, 1e-30))
            out["max_rel_loss_err"] = max(out["max_rel_loss_err"], float(((a[1] - ref[1]).abs() / (ref[1].abs() + 1e-3 / 2e-2)).max()))
            out["max_rel_prio_err"] = max(out["max_rel_prio_err"], float(((a[2] - ref[2]).abs() / (ref[2].abs() + 1.0)).max()))
            out["events"] += n_ev
    out["ok"] = bool(out["events"] > 0 and out["max_rel_grad_err"] < out["tol_grad"] and out["max_rel_loss_err"] < out["tol_loss"]
                     and out["max_rel_prio_err"] < 2e-2)
    return out


# ----------------------------------------------------------------------------------------------- multi-GPU extras
def replicas_identical(env, brains, world_size):
    """Are the replicated brains bit-identical on every rank after the timed loops?  Integer checksum of the raw bits of
    eval parameters, target parameters, Adam moments, step counters and epsilons, all-reduced with MIN and MAX."""
    import torch
    import torch.distributed as dist
    parts = []
    for b in brains:
        d = b._dev
        for t in (d.params, d.target, d.adam_m, d.adam_v):
            if t is not None:
                parts.append(t.view(torch.int32).to(torch.int64).sum().reshape(1))
        parts.append(d.adam_step.to(torch.int64).reshape(1))
    parts.append(env._eps.view(torch.int64).sum().reshape(1))
    chk = torch.cat(parts)
    if world_size == 1:
        return True
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def strong_scaling_point(args, rl, PERD3QN, torch, dist, world_size, local, total_worlds=4096):
    """BASELINE.json configs[3] as stated: a FIXED total of 4096 worlds sharded over the N GPUs (512 per GPU at N = 8), same
    loop body, device-timed like `value`.  Printed as `strong`; the driver's own scaling numbers stay the weak ones."""
    torch.manual_seed(0)
    brains = [PERD3QN(exploration=0, capacity=args.capacity), PERD3QN(exploration=0, capacity=args.capacity)]
    env = rl.Environment(width=W, height=H, brains=brains, max_agents=TARGET, update_interval=500, print_results=False,
                         training=True, n_worlds=total_worlds, seed=0, device=torch.device("cuda", local), precision=args.precision)
    env.reset(); env.top_up(TARGET)
    count = torch.zeros(1, dtype=torch.int64, device=env.device)

    def body(n_epi):
        count.add_(env.world.n_agents.sum())
        env.act(n_epi); env.step(); env.learn(n_epi); env.update_env(n_epi, top_up=TARGET)

    n_epi = 1
    for _ in range(max(3, args.warmup)):
        body(n_epi); n_epi += 1
    dist.barrier(); torch.cuda.synchronize()
    count.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        body(n_epi); n_epi += 1
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=env.device)
    agents = count.clone()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(agents, op=dist.ReduceOp.SUM)
    same = replicas_identical(env, brains, world_size)
    return {"scaling": "strong", "worlds_total": total_worlds, "worlds_per_gpu": env.n_worlds, "value": int(agents) / (float(ms) / 1e3),
            "unit": UNIT, "ms_per_step": float(ms) / args.steps, "steps": args.steps, "replicas_identical": same,
            "note": "fixed 4096 worlds over N GPUs; per-step device work shrinks with N while ~60 launches + the gradient / gate "
                    "all-reduces per step do not: launch latency and all-reduce latency bound this point at large N"}


# ----------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: reinlife_b200 has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import reinlife_b200 as rl
    from reinlife_b200.Models import PERD3QN

    torch.manual_seed(0)
    brains = [PERD3QN(exploration=0, capacity=args.capacity), PERD3QN(exploration=0, capacity=args.capacity)]
    n_worlds = args.worlds_per_gpu * world_size
    env = rl.Environment(width=W, height=H, brains=brains, max_agents=TARGET, update_interval=500, print_results=False,
                         training=True, n_worlds=n_worlds, seed=0, device=torch.device("cuda", local), precision=args.precision)
    env.reset()
    env.top_up(TARGET)
    NW, C = env.n_worlds, H * W
    count = torch.zeros(1, dtype=torch.int64, device=env.device)

    def body(n_epi):
        count.add_(env.world.n_agents.sum())     # len(env.agents) at the get_action phase, summed over worlds
        env.act(n_epi)
        env.step()
        env.learn(n_epi)
        env.update_env(n_epi, top_up=TARGET)     # update_env + the saturated-world generator in one launch (rl_world_update_top_up)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_epi = 1
    for _ in range(max(3, args.warmup)):
        body(n_epi); n_epi += 1
    barrier()

    # ---- timed region 1: device-resident, asynchronous launches ("value")
    clocks = Clocks(local) if rank == 0 else None
    count.zero_()
    launches0 = env.gpu_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        body(n_epi); n_epi += 1
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=env.device)
    agents = count.clone()
    launches = env.gpu_launches - launches0
    clk = clocks.stop() if clocks else None
    if world_size > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(agents, op=dist.ReduceOp.SUM)
    ms_total, agent_steps = float(ms), int(agents)
    value = agent_steps / (ms_total / 1e3)

    # ---- timed region 2: end to end through the public API, host in the loop every step ("e2e"):
    #      pinned control block -> device (step stamp read by the stats kernel), tracker record + event counts -> host
    count.zero_()
    nt = brains[0]._dev.dims.n_train
    h2d = env.tracker.ctrl_host[0].numel() * 8
    d2h = env.tracker.nv * 8 + 4 * len(brains)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        body(n_epi)
        rec = env.tracker.ring[(env.tracker.k - 1) % env.tracker.ring_len].cpu()        # D2H + sync (tracker record of this step)
        evs = [float(b._dev.grad[nt].cpu()) for b in brains]                            # train() events of this step
        assert int(rec[len(brains) * 8 + 3]) == n_epi, "stats record is not from this step"
        n_epi += 1
    barrier()
    sec2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=env.device)
    agents2 = count.clone()
    if world_size > 1:
        dist.all_reduce(sec2, op=dist.ReduceOp.MAX)
        dist.all_reduce(agents2, op=dist.ReduceOp.SUM)
    e2e_value = int(agents2) / float(sec2)

    # ---- per-phase device times (CUDA events on the launching stream) for the roofline
    phases = {"act": 0.0, "step": 0.0, "learn": 0.0, "update": 0.0}
    n_meas = 5
    n_agents_meas = 0
    ev_meas = 0.0
    env.kernel_events = []
    for _ in range(n_meas):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        n_agents_meas += int(env.world.n_agents.sum())
        ev[0].record(); env.act(n_epi); ev[1].record(); env.step(); ev[2].record(); env.learn(n_epi); ev[3].record()
        env.update_env(n_epi, top_up=TARGET); ev[4].record()
        torch.cuda.synchronize()
        ev_meas += sum(float(b._dev.grad[nt]) for b in brains) / world_size    # grad[nt] is all-reduced: events per GPU
        for k, name in enumerate(phases):
            phases[name] += ev[k].elapsed_time(ev[k + 1]) / n_meas
        n_epi += 1
    n_avg = n_agents_meas / n_meas / NW
    ev_avg = ev_meas / n_meas
    parity = parity_check(env, brains, args.precision) if rank == 0 else None
    same = replicas_identical(env, brains, world_size)          # before parity_check's re-runs matter: they touch grad / loss only
    # the event kernel alone (one launch per brain per step): CUDA events recorded around rl_brain_learn(_tc)
    k_ms = [a.elapsed_time(b) for (nm, _, a, b) in env.kernel_events if nm == "learn_events"]
    u_ms = [a.elapsed_time(b) for (nm, _, a, b) in env.kernel_events if nm == "world_update"]
    update_kernel_ms = sum(u_ms) / max(1, len(u_ms)) if u_ms else phases["update"]
    env.kernel_events = None
    learn_kernel_ms = sum(k_ms) / max(1, len(k_ms))
    ev_per_launch = ev_avg / len(brains)
    learn_kernel_name = {"fp16": "k_learn_dueling_p", "tf32": "k_learn_dueling_tc2", "fp32": "k_learn_dueling"}[args.precision]
    traffic = None
    try:
        # static number: dram__bytes_read + dram__bytes_write of ONE launch from the committed `ncu --set full` capture of this
        # same command (profiles/ncu_full_r02c_summary.md), not measured in this run
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02c.json"))).get(learn_kernel_name + "_bytes_per_launch")
    except Exception:
        pass
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bf16_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    # algorithmic bytes per world: step 2C+30n, observe C+12n+612n
    # + 306 n when the World kernels also emit the float16 rows get_action / the replay store consume (precision="fp16", dueling brains)
    row_b = 612 + (306 if getattr(env.world, "obs_state_h", None) is not None else 0)
    b_step = NW * ((2 * C + 30 * n_avg) + (C + 12 * n_avg + row_b * n_avg))
    b_obs = NW * (C + 12 * n_avg + row_b * n_avg)
    flop_event = 2.0 * 64 * (2 * 53504 + 2 * 53504) - 2.0 * 64 * (153 * 128)   # 2 forwards + backward (dX of layer 1 not needed)
    roof_k = {
        "k_world_step": {"bound": "hbm", "ms": phases["step"], "achieved": b_step / phases["step"] / 1e6, "peak": hbm_peak, "unit": "GB/s"},
        "k_world_update": {"bound": "hbm", "ms": update_kernel_ms, "achieved": b_obs / update_kernel_ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                           "note": "update_env + the benchmark's saturated-world top-up fused in one launch (rl_world_update_top_up): one list rebuild, one "
                                   "observation pass; CUDA events around the launch (phase_ms.update also holds the tracker statistics kernel)"},
        learn_kernel_name: {"bound": "tensor", "ms": learn_kernel_ms, "launches_per_step": len(brains),
                                "achieved": ev_per_launch * flop_event / max(learn_kernel_ms, 1e-9) / 1e9, "peak": bf16_peak, "unit": "TFLOP/s",
                                "note": ("one launch = all train() events of one brain (25.6 MFLOP per 64-row event), timed alone with CUDA events; "
                                         + {"fp16": "tcgen05 kind::f16 (fp16 operands), TMEM accumulators; peak = measured sustained bf16/fp16 tensor peak",
                                            "tf32": "tcgen05 kind::tf32, TMEM accumulators; peak shown is the measured sustained bf16 tensor peak (tf32 dense is half of it)",
                                            "fp32": "fp32 FMA on CUDA cores (reference precision); peak shown is the bf16 tensor peak"}[args.precision])},
        "act(get_action kernels x2 + row lists)": {"bound": "tensor", "ms": phases["act"], "achieved": NW * n_avg * 107008 / phases["act"] / 1e9,
                                                   "peak": bf16_peak, "unit": "TFLOP/s",
                                                   "note": {"fp16": "k_act_dueling_p, tcgen05 kind::f16", "tf32": "k_act_dueling_tc, tcgen05 kind::tf32",
                                                            "fp32": "k_brain_act, fp32 FMA on CUDA cores"}[args.precision]},
    }
    for v in roof_k.values():
        v["frac"] = v["achieved"] / v["peak"]
    dominant = max(roof_k, key=lambda k: roof_k[k]["ms"] * roof_k[k].get("launches_per_step", 1))
    roofline = dict(roof_k[dominant], kernel=dominant, traffic=traffic if dominant == learn_kernel_name else None,
                    traffic_source="profiles/traffic_r02c.json (static: one launch under ncu --set full, same command)", peak_source=peak_src,
                    step_share=roof_k[dominant]["ms"] * roof_k[dominant].get("launches_per_step", 1) / sum(phases.values()))

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp16": "fp16 operands (11 significant bits, like tf32) in get_action and the train() events, fp32 accumulate; world state exact integers, Adam fp32",
                      "tf32": "tf32 (tensor-core forward/backward products, fp32 accumulate; world state exact integers, Adam fp32)", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": workload_name(args), "grid": [H, W], "worlds_total": n_worlds, "agents_per_world": n_avg,
                       "train_events_per_step_per_gpu": ev_avg, "parallelism": f"worlds sharded x{world_size}, brains replicated, "
                       "1 NCCL all-reduce of summed gradients per step" if world_size > 1 else "single GPU",
                       "l2": "per-step working set (2 x 262 MB float32 + 2 x 131 MB float16 observation tensors + replay rows) exceeds the 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "public Environment API; per step: pinned control block H2D, tracker record + event counts D2H (host sync)"},
            "gpu_launches": launches, "clocks": clk, "roofline": roofline, "roofline_kernels": roof_k,
            "phase_ms": phases, "parity_check": parity, "replicas_identical": same}
    if world_size > 1 and not args.no_strong:
        line["strong"] = strong_scaling_point(args, rl, PERD3QN, torch, dist, world_size, local)
    if rank == 0:
        if world_size == 1 and not args.no_cpu_baseline:
            try:
                val, procs, sample, _ = cpu_arm(args, args.cpu_steps, 2)
                port = {"value": val, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample}
            except Exception as e:   # the baseline is a report, never a reason to lose the GPU number
                port = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
            try:
                ref = ref_python_arm(args.ref_steps, 5)
            except Exception as e:
                ref = None
                port["reference_error"] = str(e)[-300:]
            if ref is not None:              # the unmodified Python reference is the baseline; the (7x faster) oracle port beside it
                line["cpu_baseline"] = {"value": ref[0], "unit": UNIT, "cores": ref[1], "kind": "reference", "sample": ref[2]}
                line["cpu_baseline_port"] = port
            else:
                line["cpu_baseline"] = port
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()


def run_secondary(args):
    """Secondary workloads (single GPU, device-timed with CUDA events, same metric): printed as ONE JSON line each."""
    import torch
    import reinlife_b200 as rl
    from reinlife_b200.Models import DQN, PERD3QN, PPO
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: reinlife_b200 has no CPU fallback")
    torch.manual_seed(0)
    if args.workload == "c2":
        import numpy as np
        z = np.load(os.path.join(ROOT, "tests", "golden", "brain_golden.npz"))
        b = DQN(training=False)
        b.agent.load_state_dict({k[4:]: z[k] for k in z.files if k.startswith("dqn/")})      # pretrained/DQN/DQN/brain_gene_0.pt
        env = rl.Environment(width=30, height=30, brains=[b], max_agents=100, print_results=False, training=False, n_worlds=256, seed=0)
        env.reset(); env.top_up(100)
        name = "BASELINE configs[1]: 256 worlds, 30x30, saturated to 100 agents, DQN x1 pretrained weights, inference only (tester loop body)"

        def body(n_epi):
            env.act(n_epi); env.step(); env.update_env(n_epi, top_up=100)
        warm, steps = 20, max(args.steps, 100)
    else:
        brains = [PPO(), PERD3QN(exploration=20, capacity=1000)]
        env = rl.Environment(width=60, height=60, brains=brains, max_agents=400, print_results=False, training=True,
                             static_families=False, n_worlds=4, seed=0, slot_cap=1024, update_interval=10 ** 9)
        env.reset()
        name = ("BASELINE configs[4] brain mix and grid (60x60, max_agents=400, [PPO, PERD3QN], static_families=False) at REDUCED scale: "
                "4 natural worlds, exact per-lineage brain pools driven per agent from the host (the batched pool for 1024 worlds is not built)")

        def body(n_epi):
            env.act(n_epi); env.step(); env.learn(n_epi); env.update_env(n_epi)
        warm, steps = 150, max(args.steps, 50)
    count = 0
    n_epi = 1
    for _ in range(warm):
        body(n_epi); n_epi += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cnt = torch.zeros(1, dtype=torch.int64, device=env.device)
    e0.record()
    for _ in range(steps):
        cnt.add_(env.world.n_agents.sum())
        body(n_epi); n_epi += 1
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    env.check_status()
    print(json.dumps({"metric": METRIC, "value": int(cnt) / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warm,
                      "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "fp32" if args.workload == "c5" else "fp32 (DQN forward on CUDA cores)", "data": "synthetic",
                      "config": {"workload": name, "agents_per_world": int(cnt) / steps / env.n_worlds,
                                 "max_gene": getattr(env, "max_gene", None)}, "secondary": True}))


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "c3":
        run_secondary(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

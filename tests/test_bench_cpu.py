"""bench.py contract checks that need no GPU: the reference arm (the unmodified Python reference from baseline/_ref when
it is installed, else the oracle port of the same loop, on the host cores) prints ONE
JSON line with the keys the driver reads, and the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0", "--cpu-worlds-per-core", "1"], cwd=ROOT, text=True, timeout=600)
    lines = [ln for ln in out.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent_steps_per_sec" and d["unit"] == "agent*step/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and "workload" in d["config"]
    cb, e2e = d["cpu_baseline"], d["e2e"]
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "ReinLife"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)

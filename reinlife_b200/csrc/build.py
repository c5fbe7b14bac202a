"""Builds libreinlife_b200.so in-tree with nvcc for sm_100a (no torch headers: the boundary is a C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["c_api.cu", "world_kernels.cu", "stats_kernels.cu", "rows_kernels.cu", "brain_kernels.cu", "replay_kernels.cu", "learn_kernels.cu", "learn_rows_kernels.cu", "ppo_kernels.cu", "sumtree_kernels.cu", "tc_selftest.cu", "tc_issue_probe.cu", "tc_kernels.cu", "tc_pair_kernels.cu", "tc_act_kernels.cu", "tc_dqn_kernels.cu"]
OUT = os.path.join(os.path.dirname(HERE), "libreinlife_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("RL_NVCC_EXTRA", "").split()   # e.g. -DRL_WT=128 (experiments)


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh"))]
    deps += [os.path.join(HERE, "..", "..", "include", f) for f in ("reinlife_b200.h", "rl_rng.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    objs = []
    procs = []
    for s in srcs:
        o = s[:-3] + ".o"
        objs.append(o)
        procs.append((s, subprocess.Popen([NVCC, *FLAGS, "-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    ok = True
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            ok = False
            sys.stderr.write(out)
        elif verbose:
            sys.stderr.write(out)
    if not ok:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([NVCC, "-shared", "-Wno-deprecated-gpu-targets", "-o", OUT, *objs, "-lcudart"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares; host logic; RNG spec."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import numpy as np
import pytest


def test_library_exports_every_declared_symbol():
    from reinlife_b200 import _lib
    lib = _lib.load()
    syms = _lib.exported_symbols()
    assert len(syms) >= 20 and "rl_world_step" in syms and "rl_brain_learn" in syms
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.rl_version() >= 100


def test_argument_errors_are_reported_without_a_gpu():
    from reinlife_b200 import _lib
    lib = _lib.load()
    cfg = _lib.WorldCfg(1, 2, 30, 2, 100, 60, 160, 1, 0, 1, 0, 0)      # height 2 < 3  (World/grid.py:23-24)
    bufs = _lib.WorldBufs()
    rc = lib.rl_world_step(C.byref(cfg), C.byref(bufs), C.c_uint64(1), None)
    assert rc == -1 and b"height" in lib.rl_last_error()
    with pytest.raises(_lib.RLError):
        _lib.check(rc)
    d = __import__("reinlife_b200.Models.packing", fromlist=["x"]).dims(0)
    assert (d.n1, d.n2, d.nh) == (128, 256, 9) and d.n_train % 4 == 0 and d.n_total == d.off_w2 + 128 * 256


def test_struct_layouts_match_the_header():
    """ctypes mirrors == what a C compiler makes of include/reinlife_b200.h (sizes and the offset of the last member)."""
    import subprocess, tempfile
    from reinlife_b200 import _lib
    from reinlife_b200.World.vecworld import REC_DTYPE
    assert REC_DTYPE.itemsize == 16
    pairs = [("rl_world_cfg", _lib.WorldCfg, "world_id0"), ("rl_world_bufs", _lib.WorldBufs, "obs_prime_h"),
             ("rl_rows_bufs", _lib.RowsBufs, "row_cap"), ("rl_replay_bufs", _lib.ReplayBufs, "obs_fp16"),
             ("rl_learn_bufs", _lib.LearnBufs, "lr"), ("rl_brain_act", _lib.BrainAct, "epsilon"),
             ("rl_brain_sched", _lib.BrainSched, "max_epi"), ("rl_ppo_bufs", _lib.PpoBufs, "eps_clip"),
             ("rl_sumtree_bufs", _lib.SumTreeBufs, "p_new"), ("rl_ns_best", _lib.NsBest, "brain"),
             ("rl_ns_state", _lib.NsState, "best"), ("rl_world_ns_bufs", _lib.WorldNsBufs, "n_lineages"),
             ("rl_agent_rec", None, "prev_slot")]
    body = "".join(f'printf("%s %zu %zu\\n", "{c}", sizeof({c}), offsetof({c}, {m}));' for c, _, m in pairs)
    src = f'#include <stdio.h>\n#include <stddef.h>\n#include "{os.path.join(ROOT, "include", "reinlife_b200.h")}"\nint main(void){{{body}return 0;}}\n'
    with tempfile.TemporaryDirectory() as td:
        cfile, exe = os.path.join(td, "abi.c"), os.path.join(td, "abi")
        open(cfile, "w").write(src)
        subprocess.check_call(["gcc", "-o", exe, cfile])
        out = subprocess.check_output([exe], text=True)
    got = {ln.split()[0]: (int(ln.split()[1]), int(ln.split()[2])) for ln in out.strip().splitlines()}
    for cname, ct, member in pairs:
        if ct is None:
            assert got[cname] == (16, 14)
            continue
        assert C.sizeof(ct) == got[cname][0], cname
        assert getattr(ct, member).offset == got[cname][1], (cname, member)


def test_pack_unpack_round_trip_and_reference_key_names():
    from reinlife_b200.Models import packing
    from brain_golden_util import state_dict
    for kind, name in ((0, "perd3qn"), (1, "dqn"), (2, "ppo")):
        sd = state_dict(name)
        back = packing.unpack(kind, packing.pack(kind, sd))
        assert list(back.keys()) == list(sd.keys())
        for k in sd:
            assert (back[k].numpy() == sd[k]).all(), (name, k)
        m = packing.grad_mask(kind)
        assert int(m.sum()) == sum(v.size for v in sd.values())       # trainable entries == reference parameter count


def test_rng_spec_c_python_numpy_agree():
    """include/rl_rng.h (through the C oracle build), its python restatement and the numpy one are the same function."""
    from oracle import ref_harness as rh
    from oracle import cpu_port
    src = r'''
    #include "../include/rl_rng.h"
    unsigned long long t_draw(unsigned long long seed, unsigned long long w, unsigned long long step, unsigned site, unsigned idx) {
        return rl_draw(rl_world_key(seed, w), step, site, idx); }
    double t_uni(unsigned long long b) { return rl_uniform(b); }
    unsigned t_below(unsigned long long b, unsigned n) { return rl_below(b, n); }
    '''
    import subprocess, tempfile
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle")
    with tempfile.TemporaryDirectory() as td:
        cfile = os.path.join(here, "_rng_test.c")
        open(cfile, "w").write(src)
        so = os.path.join(td, "rng.so")
        try:
            subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", so, cfile])
        finally:
            os.remove(cfile)
        lib = C.CDLL(so)
        lib.t_draw.restype = C.c_uint64; lib.t_uni.restype = C.c_double; lib.t_below.restype = C.c_uint32
        lib.t_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        lib.t_uni.argtypes = [C.c_uint64]; lib.t_below.argtypes = [C.c_uint64, C.c_uint32]
        rng = np.random.default_rng(0)
        for _ in range(200):
            seed, w, step = int(rng.integers(1 << 62)), int(rng.integers(1 << 40)), int(rng.integers(1 << 30))
            site, idx = int(rng.integers(1, 31)), int(rng.integers(1 << 20))
            c = lib.t_draw(seed, w, step, site, idx)
            p = rh.draw(rh.world_key(seed, w), step, site, idx)
            n = int(cpu_port.draws(cpu_port.world_keys(seed, [w]), step, site, [idx])[0])
            assert c == p == n
            assert lib.t_uni(c) == rh.uniform(p) == float(cpu_port.uniform(np.array([n], np.uint64))[0])
            assert lib.t_below(c, 37) == rh.below(p, 37) == int(cpu_port.below(np.array([n], np.uint64), 37)[0])


def test_no_product_code_touches_the_oracle():
    """reinlife_b200/ may mention the oracle in comments but must never import, link or load it."""
    import re
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "reinlife_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle|librl_oracle|#include\s+\"[^\"]*oracle|import_module\(.oracle", re.M)
    n = 0
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                n += 1
                assert not bad.search(open(os.path.join(dp, f)).read()), (dp, f)
    assert n > 15

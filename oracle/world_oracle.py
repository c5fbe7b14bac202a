"""TEST INFRASTRUCTURE -- ctypes front-end of oracle/rl_oracle.c (the C restatement of
ReinLife/World/environment.py; see that file's header).  Never imported by reinlife_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REC_DTYPE = np.dtype([("cell", "<u2"), ("health", "<i2"), ("age", "<i2"), ("max_age", "<i2"),
                      ("gene", "<i4"), ("flags", "u1"), ("action", "i1"), ("prev_slot", "<u2")])
OBS_DIM = 153


class Cfg(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32), ("n_genes", C.c_int32), ("max_agents", C.c_int32),
                ("static_families", C.c_int32), ("limit_reproduction", C.c_int32),
                ("incentivize_killing", C.c_int32), ("_pad", C.c_int32), ("seed", C.c_uint64)]


class Best(C.Structure):
    _fields_ = [("serial", C.c_int64), ("fitness", C.c_double), ("brain", C.c_int32), ("_pad", C.c_int32)]


class NsState(C.Structure):
    """rlo_ns: what static_families=False adds to one world (max_gene, the last _produce, the ten best agents)."""
    _fields_ = [("max_gene", C.c_int32), ("produced_gene", C.c_int32), ("produced_src_best", C.c_int32),
                ("produced_src_brain", C.c_int32), ("next_serial", C.c_int64), ("best", Best * 10)]


def build(force=False):
    so = os.path.join(_HERE, "librl_oracle.so")
    src = os.path.join(_HERE, "rl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "librl_oracle.so"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleWorlds:
    """N independent worlds on the host, same buffer layout as rl_world_bufs (slot_cap rows per world)."""

    def __init__(self, n_worlds, height=30, width=30, n_genes=2, max_agents=100, seed=0, world_id0=0,
                 slot_cap=None, limit_reproduction=False, incentivize_killing=True, static_families=True):
        self.n_worlds, self.H, self.W = n_worlds, height, width
        self.C = height * width
        self.S = slot_cap or self.C
        self.world_id0 = world_id0
        self.cfg = Cfg(height, width, n_genes, max_agents, int(static_families), int(limit_reproduction),
                       int(incentivize_killing), 0, seed)
        self.static = bool(static_families)
        self.type = np.zeros((n_worlds, self.C), np.uint8)
        self.rec = np.zeros((n_worlds, self.S), REC_DTYPE)
        self.n = np.zeros(n_worlds, np.int32)
        self.reward = np.zeros((n_worlds, self.S), np.float64)
        self.obs = np.zeros((n_worlds, self.S, OBS_DIM), np.float64)
        self.t = 0
        if not self.static:                      # Agent.fitness, object identity, best_agents, max_gene (rl_oracle.c, rlo_ns)
            self.fitness = np.zeros((n_worlds, self.S), np.float64)
            self.serial = np.zeros((n_worlds, self.S), np.int64)
            self.ns = (NsState * n_worlds)()

    def _ns_args(self, w):
        return _p(self.fitness[w]), _p(self.serial[w]), C.byref(self.ns[w])

    def reset(self):
        if not self.static:
            for w in range(self.n_worlds):
                lib().rlo_reset_ns(C.byref(self.cfg), C.c_int64(self.world_id0 + w), _p(self.type[w]), _p(self.rec[w]),
                                   _p(self.n[w:w + 1]), _p(self.obs[w]), *self._ns_args(w))
            self.t = 0
            return
        lib().rlo_reset_many(C.byref(self.cfg), C.c_int64(self.world_id0), self.n_worlds, self.S,
                             _p(self.type), _p(self.rec), _p(self.n), _p(self.obs))
        self.t = 0

    def set_actions(self, actions):
        """actions: [n_worlds, S] int8 (only the first n[w] of each row are used)."""
        self.rec["action"][:, :actions.shape[1]] = actions

    def step(self):
        self.t += 1
        if not self.static:
            for w in range(self.n_worlds):
                lib().rlo_step_ns(C.byref(self.cfg), C.c_int64(self.world_id0 + w), C.c_uint64(self.t), _p(self.type[w]),
                                  _p(self.rec[w]), _p(self.n[w:w + 1]), _p(self.reward[w]), _p(self.obs[w]), *self._ns_args(w))
            return
        lib().rlo_step_many(C.byref(self.cfg), C.c_int64(self.world_id0), self.n_worlds, self.S, C.c_uint64(self.t),
                            _p(self.type), _p(self.rec), _p(self.n), _p(self.reward), _p(self.obs))

    def update(self):
        if not self.static:
            for w in range(self.n_worlds):
                lib().rlo_update_ns(C.byref(self.cfg), C.c_int64(self.world_id0 + w), C.c_uint64(self.t), _p(self.type[w]),
                                    _p(self.rec[w]), _p(self.n[w:w + 1]), _p(self.obs[w]), *self._ns_args(w))
            return
        lib().rlo_update_many(C.byref(self.cfg), C.c_int64(self.world_id0), self.n_worlds, self.S, C.c_uint64(self.t),
                              _p(self.type), _p(self.rec), _p(self.n), _p(self.obs))

    def top_up(self, target, max_age=50):
        if not self.static:
            raise NotImplementedError("the saturated-world generator is defined for static families only")
        lib().rlo_topup_many(C.byref(self.cfg), C.c_int64(self.world_id0), self.n_worlds, self.S, C.c_uint64(self.t),
                             target, max_age, _p(self.type), _p(self.rec), _p(self.n), _p(self.obs))

    def observe(self):
        for w in range(self.n_worlds):
            lib().rlo_observe(C.byref(self.cfg), _p(self.type[w]), _p(self.rec[w]), int(self.n[w]), _p(self.obs[w]))

    def load(self, w, typ, rec):
        self.type[w] = np.asarray(typ, np.uint8).reshape(-1)
        self.n[w] = len(rec)
        self.rec[w, :len(rec)] = rec

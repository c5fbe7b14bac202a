"""ctypes binding of libreinlife_b200.so (include/reinlife_b200.h).

The library is the product: if it is missing or fails to load, importing this module raises --
there is no CPU fallback anywhere in reinlife_b200.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REINLIFE_B200_LIB") or os.path.join(_HERE, "libreinlife_b200.so")   # override: A/B builds

OBS_DIM = 153
N_ACTIONS = 8
OBS_LD = 160           # floats per observation row (640 B: 16-byte vector / TMA friendly); pad is zero
MAX_GENES = 32
N_STATS = 8

F_KILLED, F_INTER_KILLED, F_INTRA_KILLED, F_ATE_SUPER, F_REPRODUCED, F_DEAD = 1, 2, 4, 8, 16, 32
EMPTY, FOOD, POISON, AGENT, KIN, SUPER_FOOD = 0, 1, 2, 3, 4, 5


class RLError(RuntimeError):
    pass


class WorldCfg(C.Structure):
    _fields_ = [("n_worlds", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("n_genes", C.c_int32),
                ("max_agents", C.c_int32), ("slot_cap", C.c_int32), ("obs_ld", C.c_int32),
                ("static_families", C.c_int32), ("limit_reproduction", C.c_int32),
                ("incentivize_killing", C.c_int32), ("seed", C.c_uint64), ("world_id0", C.c_int64)]


class WorldBufs(C.Structure):
    _fields_ = [("type", C.c_void_p), ("rec", C.c_void_p), ("n_agents", C.c_void_p), ("reward", C.c_void_p),
                ("obs_state", C.c_void_p), ("obs_prime", C.c_void_p), ("gene_count", C.c_void_p),
                ("status", C.c_void_p), ("stats", C.c_void_p), ("reward_div100", C.c_void_p),
                ("obs_state_h", C.c_void_p), ("obs_prime_h", C.c_void_p)]


class NsBest(C.Structure):          # rl_ns_best
    _fields_ = [("serial", C.c_int64), ("fitness", C.c_double), ("brain", C.c_int32), ("_pad", C.c_int32)]


class NsState(C.Structure):         # rl_ns_state: max_gene, the last _produce event, the ten best agents (non-static families)
    _fields_ = [("max_gene", C.c_int32), ("produced_gene", C.c_int32), ("produced_src_best", C.c_int32),
                ("produced_src_brain", C.c_int32), ("next_serial", C.c_int64), ("best", NsBest * 10)]


class WorldNsBufs(C.Structure):     # rl_world_ns_bufs
    _fields_ = [("fitness", C.c_void_p), ("serial", C.c_void_p), ("state", C.c_void_p), ("n_lineages", C.c_void_p)]


_lib = None


def load():
    """Load the shared library (building it is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RLError(f"{LIB_PATH} not found: build it with `python -m reinlife_b200.csrc.build` "
                      "(nvcc, sm_100a).  reinlife_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.rl_last_error.restype = C.c_char_p
    lib.rl_version.restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RLError(f"libreinlife_b200 error {rc}: {load().rl_last_error().decode()}")


def exported_symbols():
    """Every extern "C" symbol include/reinlife_b200.h declares (used by the CPU-side ABI test)."""
    import re
    hdr = os.path.join(_HERE, "..", "include", "reinlife_b200.h")
    text = open(hdr).read()
    return sorted(set(re.findall(r"^\s*(?:int|const char\*)\s+(rl_[a-z0-9_]+)\s*\(", text, re.M)))


ROWS_ALL, ROWS_STORE, ROWS_EVENT, N_ROW_KINDS = 0, 1, 2, 3
MODEL_DUELING, MODEL_DQN, MODEL_PPO = 0, 1, 2
ACT_DUELING, ACT_DQN, ACT_PPO, ACT_PERDQN = 0, 1, 2, 3


class RowsBufs(C.Structure):
    _fields_ = [("count", C.c_void_p), ("offset", C.c_void_p), ("total", C.c_void_p), ("rows", C.c_void_p),
                ("row_cap", C.c_int32), ("_pad", C.c_int32)]


class BrainAct(C.Structure):
    _fields_ = [("kind", C.c_int32), ("rule", C.c_int32), ("params", C.c_void_p), ("epsilon", C.c_void_p)]


class BrainSched(C.Structure):
    _fields_ = [("rule", C.c_int32), ("training", C.c_int32), ("eps_min", C.c_double), ("decay", C.c_double),
                ("max_epi", C.c_int64)]


class ReplayBufs(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("next_obs", C.c_void_p), ("action", C.c_void_p), ("reward", C.c_void_p),
                ("done", C.c_void_p), ("prio", C.c_void_p), ("pw", C.c_void_p), ("len", C.c_void_p), ("pos", C.c_void_p),
                ("maxst", C.c_void_p),
                ("capacity", C.c_int32), ("prioritized", C.c_int32), ("obs_fp16", C.c_int32), ("_pad", C.c_int32)]


class LearnBufs(C.Structure):
    _fields_ = [("params", C.c_void_p), ("target", C.c_void_p), ("grad_scratch", C.c_void_p), ("grad", C.c_void_p),
                ("adam_m", C.c_void_p), ("adam_v", C.c_void_p), ("mask", C.c_void_p), ("adam_step", C.c_void_p),
                ("new_prio", C.c_void_p), ("loss", C.c_void_p), ("kind", C.c_int32), ("batch", C.c_int32),
                ("gamma", C.c_float), ("lr", C.c_float)]


class PpoBufs(C.Structure):
    _fields_ = [("traj", ReplayBufs), ("seg_end", C.c_void_p), ("n_cons", C.c_void_p), ("row_off", C.c_void_p),
                ("flat_src", C.c_void_p), ("row_T", C.c_void_p), ("row_end", C.c_void_p), ("td", C.c_void_p),
                ("delta", C.c_void_p), ("adv", C.c_void_p), ("status", C.c_void_p), ("row_cap", C.c_int32),
                ("lmbda", C.c_float), ("eps_clip", C.c_float), ("_pad", C.c_int32)]


class SumTreeBufs(C.Structure):
    _fields_ = [("tree", C.c_void_p), ("beta", C.c_void_p), ("status", C.c_void_p), ("capacity", C.c_int32),
                ("train_start", C.c_int32), ("p_new", C.c_float), ("_pad", C.c_int32)]

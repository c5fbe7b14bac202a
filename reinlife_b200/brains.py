"""Device state of one brain (network, optimizer, per-world replay rings) + the learn step driven through the C ABI.

This is the batched stand-in for the object a reference brain holds: eval/target nn.Modules, torch.optim.Adam and
a replay buffer (Models/PERD3QN.py:50-55, Models/D3QN.py:57-62, Models/DQN.py:48-52, Models/PPO.py:96-99).
"""
import ctypes as C

import torch

from . import _lib
from .Models import packing


class DeviceBrain:
    def __init__(self, kind, state_dict, device, lr=1e-3, gamma=0.99, batch=64, has_target=True):
        self.host_kind, self.kind, self.device = kind, packing.device_kind(kind), torch.device(device)   # kind = rl_model_kind
        self.lib = _lib.load()
        self.dims = packing.dims(kind)
        self.lr, self.gamma, self.batch = float(lr), float(gamma), int(batch)
        flat = torch.from_numpy(packing.pack(kind, state_dict))      # host kind: PERDQN packs into the padded DQN layout
        self.params = flat.to(self.device)
        self.target = self.params.clone() if has_target else None
        nt = self.dims.n_train
        self.adam_m = torch.zeros(nt, device=self.device)
        self.adam_v = torch.zeros(nt, device=self.device)
        self.mask = torch.from_numpy(packing.grad_mask(kind)).to(self.device)
        self.grad = torch.zeros(nt + 4, device=self.device)
        self.adam_step = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.grad_scratch = None
        self.new_prio = self.loss = self.sample_idx = None
        self.learn_bufs = None
        self.wimg_e = self.wimg_t = None          # tensor-core weight images (tf32)
        self.wimg_eh = self.wimg_th = None        # fp16 weight images (precision="fp16" event kernel)
        self.use_fp16 = False
        self.wimg_stale = True

    # -- tensor-core weight images -------------------------------------------------------------------
    def build_wimg(self, stream_ptr, which="both"):
        """Refresh the tcgen05 operand images from the kernel-layout parameters (after Adam / load / target sync)."""
        if self.wimg_e is None:
            n = self.lib.rl_tc_wimg_floats()
            self.wimg_e = torch.zeros(n, device=self.device)
            self.wimg_t = torch.zeros(n, device=self.device)
            self.wimg_eh = torch.zeros(n, dtype=torch.float16, device=self.device)
            self.wimg_th = torch.zeros(n, dtype=torch.float16, device=self.device)
        with torch.cuda.device(self.device):
            if self.use_fp16:
                if which in ("both", "eval"):
                    _lib.check(self.lib.rl_brain_build_wimg_h(C.c_int32(self.kind), C.c_void_p(self.params.data_ptr()),
                                                              C.c_void_p(self.wimg_eh.data_ptr()), stream_ptr))
                if which in ("both", "target") and self.target is not None:
                    _lib.check(self.lib.rl_brain_build_wimg_h(C.c_int32(self.kind), C.c_void_p(self.target.data_ptr()),
                                                              C.c_void_p(self.wimg_th.data_ptr()), stream_ptr))
                return                                      # fp16 mode: get_action and the events read the fp16 images only
            if which in ("both", "eval"):
                _lib.check(self.lib.rl_brain_build_wimg(C.c_int32(self.kind), C.c_void_p(self.params.data_ptr()),
                                                        C.c_void_p(self.wimg_e.data_ptr()), stream_ptr))
            if which in ("both", "target") and self.target is not None:
                _lib.check(self.lib.rl_brain_build_wimg(C.c_int32(self.kind), C.c_void_p(self.target.data_ptr()),
                                                        C.c_void_p(self.wimg_t.data_ptr()), stream_ptr))

    # -- state_dict round trip (reference key names) -----------------------------------------------
    def state_dict(self, target=False):
        return packing.unpack(self.host_kind, (self.target if target else self.params).cpu().numpy())

    def load_state_dict(self, sd, target=False):
        flat = torch.from_numpy(packing.pack(self.host_kind, sd)).to(self.device)
        (self.target if target else self.params).copy_(flat)
        self.wimg_stale = True

    # -- learn buffers -----------------------------------------------------------------------------
    def alloc_learn(self, row_cap, need_batch_bufs=True):
        with torch.cuda.device(self.device):
            n_cta = self.lib.rl_learn_grid()
        nt = self.dims.n_train
        self.grad_scratch = torch.zeros((n_cta, nt), device=self.device)
        nb = row_cap if need_batch_bufs else 1          # PPO has no sampled batches / priorities
        self.new_prio = torch.zeros((nb, self.batch), device=self.device)
        self.loss = torch.zeros(row_cap, device=self.device)
        self.sample_idx = torch.zeros((nb, self.batch), dtype=torch.int32, device=self.device)
        self.learn_bufs = _lib.LearnBufs(self.params.data_ptr(), self.target.data_ptr() if self.target is not None else None,
                                         self.grad_scratch.data_ptr(), self.grad.data_ptr(), self.adam_m.data_ptr(),
                                         self.adam_v.data_ptr(), self.mask.data_ptr(), self.adam_step.data_ptr(),
                                         self.new_prio.data_ptr(), self.loss.data_ptr(), self.kind, self.batch,
                                         self.gamma, self.lr)

    def act_desc(self, rule, eps_ptr):
        return _lib.BrainAct(self.kind, rule, self.params.data_ptr(), eps_ptr)


class ReplayRings:
    """One ring per local world for one brain (rl_replay_bufs)."""

    def __init__(self, n_worlds, capacity, device, prioritized=True, ld=_lib.OBS_LD, fp16=False):
        dev = torch.device(device)
        self.n_worlds, self.capacity, self.prioritized, self.fp16 = n_worlds, int(capacity), bool(prioritized), bool(fp16)
        # fp16=True: rows rounded to float16 once, at store time (rl_replay_bufs.obs_fp16) -- the ring of the dueling brains
        # under precision="fp16", whose tensor-core event kernel consumes fp16 operands anyway
        dt = torch.float16 if fp16 else torch.float32
        self.obs = torch.empty((n_worlds, capacity, ld), dtype=dt, device=dev)
        self.next_obs = torch.empty((n_worlds, capacity, ld), dtype=dt, device=dev)
        self.action = torch.zeros((n_worlds, capacity), dtype=torch.int8, device=dev)
        self.reward = torch.zeros((n_worlds, capacity), device=dev)
        self.done = torch.zeros((n_worlds, capacity), dtype=torch.uint8, device=dev)
        self.prio = torch.zeros((n_worlds, capacity), device=dev)
        self.pw = torch.zeros((n_worlds, capacity), device=dev)
        self.len = torch.zeros(n_worlds, dtype=torch.int32, device=dev)
        self.pos = torch.zeros(n_worlds, dtype=torch.int32, device=dev)
        # {max(priorities) as float bits, count of entries holding it} per ring, kept exact by the store / priority-update kernels
        # (count 0 = unknown -> the next store scans the ring); zero it after writing prio[] by hand
        self.maxst = torch.zeros((n_worlds, 2), dtype=torch.int32, device=dev)
        self.bufs = _lib.ReplayBufs(self.obs.data_ptr(), self.next_obs.data_ptr(), self.action.data_ptr(),
                                    self.reward.data_ptr(), self.done.data_ptr(), self.prio.data_ptr(), self.pw.data_ptr(),
                                    self.len.data_ptr(), self.pos.data_ptr(), self.maxst.data_ptr(), self.capacity, int(self.prioritized),
                                    int(self.fp16), 0)

    @staticmethod
    def bytes_needed(n_worlds, capacity, ld=_lib.OBS_LD, fp16=False):
        return n_worlds * capacity * (2 * ld * (2 if fp16 else 4) + 1 + 4 + 1 + 4 + 4)


class SumTrees:
    """PERDQN's Memory per local world (rl_sumtree_bufs): float64 sum trees + beta; the transitions themselves live in
    the brain's ReplayRings (pos = SumTree.write, len = n_entries)."""

    def __init__(self, n_worlds, capacity, device, train_start=1000, ev_cap=1):
        dev = torch.device(device)
        self.capacity = int(capacity)
        self.tree = torch.zeros((n_worlds, 2 * self.capacity - 1), dtype=torch.float64, device=dev)
        self.beta = torch.full((n_worlds,), 0.4, dtype=torch.float64, device=dev)          # Memory.beta, PERDQN.py:266
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.ev_weight = torch.zeros(int(ev_cap), device=dev)
        # Memory._get_priority(0) as append_sample computes it -- float32 torch scalars (PERDQN.py:126-128,272-273)
        self.p_new = float((torch.zeros((), dtype=torch.float32).abs() + 0.01) ** 0.6)
        self.bufs = _lib.SumTreeBufs(self.tree.data_ptr(), self.beta.data_ptr(), self.status.data_ptr(), self.capacity,
                                     int(train_start), self.p_new, 0)

    @staticmethod
    def bytes_needed(n_worlds, capacity):
        return n_worlds * (2 * capacity - 1) * 8


class PpoData:
    """PPO's per-(world, brain) data list + the per-step segment plan (rl_ppo_bufs)."""

    def __init__(self, n_worlds, capacity, row_cap, device, lmbda=0.95, eps_clip=0.1):
        dev = torch.device(device)
        self.traj = ReplayRings(n_worlds, capacity, dev, prioritized=True)        # prio[] = pi_old(a)
        self.seg_end = torch.zeros((n_worlds, capacity), dtype=torch.uint8, device=dev)
        self.n_cons = torch.zeros(n_worlds, dtype=torch.int32, device=dev)
        self.row_off = torch.zeros(n_worlds + 1, dtype=torch.int32, device=dev)
        self.row_cap = int(row_cap)
        self.flat_src = torch.zeros(self.row_cap, dtype=torch.int32, device=dev)
        self.row_T = torch.ones(self.row_cap, dtype=torch.int32, device=dev)
        self.row_end = torch.zeros(self.row_cap, dtype=torch.uint8, device=dev)
        self.td = torch.zeros(self.row_cap, device=dev)
        self.delta = torch.zeros(self.row_cap, device=dev)
        self.adv = torch.zeros(self.row_cap, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.bufs = _lib.PpoBufs(self.traj.bufs, self.seg_end.data_ptr(), self.n_cons.data_ptr(), self.row_off.data_ptr(),
                                 self.flat_src.data_ptr(), self.row_T.data_ptr(), self.row_end.data_ptr(),
                                 self.td.data_ptr(), self.delta.data_ptr(), self.adv.data_ptr(), self.status.data_ptr(),
                                 self.row_cap, float(lmbda), float(eps_clip), 0)

    # what the generic code reads from a brain's `_replay`
    @property
    def len(self):
        return self.traj.len


def learn_step(world, rows, gene, brain, replay, t, allreduce=None):
    """store -> sample -> train events -> (all-reduce) -> Adam -> priorities, for one brain.  No host sync."""
    lib, st = world.lib, world._stream()
    with torch.cuda.device(world.device):
        _lib.check(lib.rl_replay_store(C.byref(world.cfg), C.byref(world.bufs), C.byref(rows.bufs), C.c_int32(gene),
                                       C.byref(replay.bufs), st))
        _lib.check(lib.rl_replay_sample(C.byref(world.cfg), C.byref(rows.bufs), C.c_int32(gene), C.byref(replay.bufs),
                                        C.c_int32(brain.batch), C.c_uint64(t), C.c_void_p(brain.sample_idx.data_ptr()), st))
        _lib.check(lib.rl_brain_learn(C.byref(world.cfg), C.byref(rows.bufs), C.c_int32(gene), C.byref(replay.bufs),
                                      C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs), st))
        if allreduce is not None:
            allreduce(brain.grad)
        _lib.check(lib.rl_brain_adam(C.byref(brain.learn_bufs), st))
        _lib.check(lib.rl_replay_update_prio(C.byref(world.cfg), C.byref(rows.bufs), C.c_int32(gene), C.byref(replay.bufs),
                                             C.c_int32(brain.batch), C.c_void_p(brain.sample_idx.data_ptr()),
                                             C.c_void_p(brain.new_prio.data_ptr()), st))


def sync_target(brain, world, cond_ptr=None):
    with torch.cuda.device(world.device):
        _lib.check(world.lib.rl_brain_sync_target(C.byref(brain.learn_bufs), C.c_void_p(cond_ptr), world._stream()))

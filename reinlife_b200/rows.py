"""Per-brain row lists (rl_rows_bufs) -- the batched stand-in for `for agent in env.agents` (Helpers/trainer.py:88-96)."""
import ctypes as C

import torch

from . import _lib


class RowLists:
    def __init__(self, world, row_cap=None):
        self.world = world
        G, NW = world.G, world.n_worlds
        self.row_cap = int(row_cap or NW * world.S)
        dev = world.device
        K = _lib.N_ROW_KINDS
        self.count = torch.zeros((G * K, NW), dtype=torch.int32, device=dev)
        self.offset = torch.zeros((G * K, NW), dtype=torch.int32, device=dev)
        self.total = torch.zeros(G * K, dtype=torch.int32, device=dev)
        self.rows = torch.zeros((G * K, self.row_cap), dtype=torch.int32, device=dev)
        self.bufs = _lib.RowsBufs(self.count.data_ptr(), self.offset.data_ptr(), self.total.data_ptr(),
                                  self.rows.data_ptr(), self.row_cap, 0)

    def build(self, kinds_mask=1, train_freq=None, event_on=None):
        w = self.world
        G = w.G
        tf = (C.c_int32 * G)(*(train_freq or [1] * G))
        on = (C.c_int32 * G)(*(event_on or [0] * G))
        with torch.cuda.device(w.device):
            _lib.check(w.lib.rl_rows_build(C.byref(w.cfg), C.byref(w.bufs), C.byref(self.bufs), tf, on,
                                           C.c_int32(kinds_mask), w._stream()))

    def list(self, gene, kind):
        """Host copy of one list (tests)."""
        i = gene * _lib.N_ROW_KINDS + kind
        n = int(self.total[i])
        return self.rows[i, :n].cpu().numpy()

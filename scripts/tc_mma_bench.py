"""Cycles per tcgen05.mma kind::tf32 by operand layout (timing probe, one CTA)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reinlife_b200 import _lib
lib = _lib.load()
torch.zeros(1, device="cuda")
out = (C.c_longlong * 2)()
for (M, N) in [(128, 64), (128, 128), (128, 256), (64, 16)]:
    for mode in (0, 10, 12, 20, 30, 31, 33):
        _lib.check(lib.rl_tc_mma_bench(M, N, 8, mode, 50, out))
        macs = M * N * 64 * (2 if mode >= 20 else 1)
        print(f"M={M} N={N} K=64 mode={mode}: issue {out[0]} cyc, complete {out[1]} cyc -> {macs / out[1]:.0f} MAC/clk")

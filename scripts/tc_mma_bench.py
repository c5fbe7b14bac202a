"""Cycles per tcgen05.mma by operand kind / layout (timing probe, one CTA).
mode: 0 tf32 no-swizzle, 2 tf32 SWIZZLE_128B, 10/12 the same issued from an elected lane in warp-uniform control flow,
20/21/22 kind::f16 (no-swizzle / padded chunk stride / SWIZZLE_128B), 30-33 1-4 concurrent tf32 issuer warps."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reinlife_b200 import _lib
lib = _lib.load()
torch.zeros(1, device="cuda")
out = (C.c_longlong * 2)()
for (M, N) in [(128, 64), (128, 128), (128, 256), (64, 16)]:
    for mode in (0, 2, 10, 12, 20, 21, 22, 30, 31, 33):
        for ks in (4, 8):
            _lib.check(lib.rl_tc_mma_bench(M, N, ks, mode, 50, out))
            k_per = 16 if 20 <= mode < 30 else 8
            macs = M * N * ks * k_per
            print(f"M={M} N={N} ksteps={ks} mode={mode}: issue {out[0]} cyc, complete {out[1]} cyc -> {macs / out[1]:.0f} MAC/clk "
                  f"({out[1] / ks:.0f} cyc/MMA)")

"""Cycles per tcgen05.mma kind::f16 (M = 128) by issue style, operand layout, N and number of issuing warps."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reinlife_b200 import _lib
lib = _lib.load()
torch.zeros(1, device="cuda")
out = (C.c_longlong * 2)()
for style in (0, 1):
    for layout in (0, 1, 2):
        for N in (16, 64, 128, 256):
            for nw in (1, 2):
                if nw == 2 and N > 256:
                    continue
                for nmma in (8, 32):
                    _lib.check(lib.rl_tc_issue_probe(N, nmma, layout, style, nw, 50, out))
                    floor = 128 * N / 256
                    print(f"style={style} layout={layout} N={N} warps={nw} nmma={nmma}: issue {out[0] / nmma:.0f} cyc/MMA, "
                          f"complete {out[1] / nmma:.0f} cyc/MMA per warp (math floor {floor:.0f})", flush=True)

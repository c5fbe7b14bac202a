"""Probe of tcgen05 kind::f16 operand conventions on this part (run on the GPU box):
K-major and MN-major fp16 operands read from the SAME interleaved no-swizzle image
    off(r, c) = (r/8)*(W*8) + (c/8)*64 + (r%8)*8 + (c%8)      [halves, W = image width]
(K-major: rows = M/N index, LBO = 128 B, SBO = W*16 B;  MN-major: rows = K index, SBO = 128 B, LBO = W*16 B),
plus issue / completion cycles of kind::f16 MMAs at N = 64 / 128 / 256.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinlife_b200 import _lib   # noqa: E402


def himg(mat):
    R, W = mat.shape
    out = np.zeros(R * W, np.float16)
    r, c = np.meshgrid(np.arange(R), np.arange(W), indexing="ij")
    off = (r >> 3) * (W * 8) + (c >> 3) * 64 + (r & 7) * 8 + (c & 7)
    out[off.reshape(-1)] = mat.reshape(-1)
    return out


def run(M, N, K, a_mn, b_mn):
    lib = _lib.load()
    rng = np.random.default_rng(M + 3 * N + 7 * K + a_mn + 2 * b_mn)
    A = rng.standard_normal((M, K)).astype(np.float16)
    B = rng.standard_normal((N, K)).astype(np.float16)
    # K-major: image of [MN][K]; MN-major: image of the transposed matrix [K][MN]
    a_img = himg(A.T.copy()) if a_mn else himg(A)
    b_img = himg(B.T.copy()) if b_mn else himg(B)
    geo = lambda mn, MN: (MN * 16, 128, 2 * MN * 16) if mn else (128, K * 16, 256)   # (lbo, sbo, kstep) bytes
    al, asb, ak = geo(a_mn, M)
    bl, bsb, bk = geo(b_mn, N)
    ta, tb = torch.from_numpy(a_img).cuda(), torch.from_numpy(b_img).cuda()
    d = torch.zeros((M, N), device="cuda")
    _lib.check(lib.rl_tc_gemm_test_h(C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(d.data_ptr()), M, N, K,
                                     a_img.size, b_img.size, al, asb, ak, bl, bsb, bk, a_mn, b_mn, None))
    torch.cuda.synchronize()
    want = A.astype(np.float64) @ B.astype(np.float64).T
    err = np.abs(d.cpu().numpy() - want).max() / np.abs(want).max()
    return err


if __name__ == "__main__":
    for (M, N, K) in ((128, 64, 64), (128, 256, 128), (128, 160, 128), (128, 16, 128), (64, 16, 256)):
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: rel err {run(M, N, K, a_mn, b_mn):.3e}", flush=True)
    lib = _lib.load()
    out = (C.c_longlong * 2)()
    for N in (64, 128, 160, 256):
        for ks in (1, 4, 8):
            _lib.check(lib.rl_tc_mma_bench(128, N, ks, 20, 50, out))
            print(f"f16 M=128 N={N} ksteps={ks}: issue {out[0]} cycles, complete {out[1]} cycles "
                  f"({out[1] / ks:.0f}/MMA; math at peak = {128 * N * 16 * 2 / 8192:.0f})", flush=True)

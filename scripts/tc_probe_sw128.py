"""Pins (on the GPU box) the SWIZZLE_128B conventions the paired event kernel relies on:
 1. K-major SWIZZLE_128B A operand ([rows][64-half K blocks], 16-byte unit ^= row & 7) against a no-swizzle B;
 2. the SAME image read MN-major as B (N = image columns, K = image rows) against a no-swizzle MN-major A;
 3. TMA tile::gather4 landing ring rows in exactly that image.
"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinlife_b200 import _lib
from tc_probe_h import himg

lib = _lib.load()


def sw128(mat):
    """[R][W] halves -> ceil(W/64) blocks of [R][128 B], 16-byte unit index ^= (row & 7)."""
    R, W = mat.shape
    nb = (W + 63) // 64
    out = np.zeros(nb * R * 64, np.float16)
    r, c = np.meshgrid(np.arange(R), np.arange(W), indexing="ij")
    off = (c // 64) * (R * 64) + r * 64 + ((((c % 64) // 8) ^ (r & 7)) * 8) + (c % 8)
    out[off.reshape(-1)] = mat.reshape(-1)
    return out


def gemm(a_img, b_img, M, N, K, ageo, bgeo, a_mn, b_mn, akblk, bkblk, alay, blay):
    ta, tb = torch.from_numpy(a_img).cuda(), torch.from_numpy(b_img).cuda()
    d = torch.zeros((M, N), device="cuda")
    _lib.check(lib.rl_tc_gemm_test_hx(C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(d.data_ptr()), M, N, K,
                                      a_img.size, b_img.size, *ageo, *bgeo, a_mn, b_mn, akblk, bkblk, alay, blay, None))
    torch.cuda.synchronize()
    return d.cpu().numpy()


rng = np.random.default_rng(3)
# 1. L1-like: D[128 b][128 n] = X[128][160] W1[128][160]^T, X in SWIZZLE_128B K-major, W1 no-swizzle K-major
M, N, K = 128, 128, 160
X = rng.standard_normal((M, K)).astype(np.float16); W = rng.standard_normal((N, K)).astype(np.float16)
got = gemm(sw128(X), himg(W), M, N, K, (16, 1024, 32), (128, K * 16, 256), 0, 0, M * 128, 0, 2, 0)
want = X.astype(np.float64) @ W.astype(np.float64).T
print("1. SW128 K-major A: rel err", np.abs(got - want).max() / np.abs(want).max(), flush=True)
# 2. dW1-like: D[128 k1][160 x] = sum_b dH1[b][k1] X[b][x]: A = dH1 image [128 b][128 k1] read MN-major (no swizzle),
#    B = X SWIZZLE_128B image [128 b][160 x] read MN-major: LBO = next 64-column block (16 KB), SBO = next 8 rows (1 KB), k-step 2 KB
Bk, Mk, Nx = 128, 128, 160
dH1 = rng.standard_normal((Bk, Mk)).astype(np.float16)
got = gemm(himg(dH1), sw128(X), Mk, Nx, Bk, (Mk * 16, 128, 2 * Mk * 16), (16384, 1024, 2048), 1, 1, 0, 0, 0, 2)
want = dH1.astype(np.float64).T @ X.astype(np.float64)
print("2. SW128 MN-major B (N=160): rel err", np.abs(got - want).max() / np.abs(want).max(), flush=True)
for Nx2 in (80, 64, 128):
    got = gemm(himg(dH1), sw128(X), Mk, Nx2, Bk, (Mk * 16, 128, 2 * Mk * 16), (16384, 1024, 2048), 1, 1, 0, 0, 0, 2)
    print(f"   N={Nx2}: rel err", np.abs(got - want[:, :Nx2]).max() / np.abs(want).max(), flush=True)
# 3. gather4
n_rows = 5000
ring = rng.standard_normal((n_rows, 160)).astype(np.float16)
idx = rng.integers(0, n_rows, 128).astype(np.int32)
tr, ti = torch.from_numpy(ring).cuda(), torch.from_numpy(idx).cuda()
for box_rows in (1, 4):
    out = torch.zeros(3 * 16384 // 2, dtype=torch.float16, device="cuda")
    try:
        _lib.check(lib.rl_tma_gather_test(C.c_void_p(tr.data_ptr()), C.c_longlong(n_rows), C.c_void_p(ti.data_ptr()), C.c_void_p(out.data_ptr()), box_rows))
    except Exception as e:
        print(f"3. gather4 box_rows={box_rows}: error {e}", flush=True)
        continue
    want = sw128(np.concatenate([ring[idx], np.zeros((128, 32), np.float16)], axis=1))
    got = out.cpu().numpy()
    print(f"3. gather4 box_rows={box_rows}: image equal = {np.array_equal(got.view(np.uint16), want.view(np.uint16))}, "
          f"mismatches {(got.view(np.uint16) != want.view(np.uint16)).sum()} of {got.size}", flush=True)

import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reinlife_b200 import _lib
lib = _lib.load()
def to_img(mat):
    R, K = mat.shape
    out = np.zeros(R * K, np.float32)
    r, c = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    off = (r >> 3) * (K * 8) + (c >> 2) * 32 + (r & 7) * 4 + (c & 3)
    out[off.reshape(-1)] = mat.reshape(-1)
    return out
def run(M, N, K, a_img, b_img, a_mn, b_mn, lbo=0, sbo=0, kstep=0, lt=0):
    d = torch.zeros((M, N), device="cuda")
    _lib.check(lib.rl_tc_gemm_test_ex(C.c_void_p(a_img.data_ptr()), C.c_void_p(b_img.data_ptr()), C.c_void_p(d.data_ptr()),
                                      M, N, K, a_mn, b_mn, lbo, sbo, kstep | (lt << 20), None))
    torch.cuda.synchronize()
    return d.cpu().numpy()
rng = np.random.default_rng(0)
# (1) A MN-major alone
M, N, K = 128, 64, 16
A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
g = run(M, N, K, torch.from_numpy(to_img(A.T.copy())).cuda(), torch.from_numpy(to_img(B)).cuda(), 1, 0)
print("A mn-major, B k-major: err", np.abs(g - A @ B.T).max(), "absmax", np.abs(g).max())
# (2) identity-A probe of B MN-major under each layout type
M, N, K = 64, 128, 8
A = np.zeros((M, K), np.float32)
for k in range(8): A[k, k] = 1.0
a_img = torch.from_numpy(to_img(A)).cuda()
b_lin = torch.arange(1, 8192 + 1, dtype=torch.float32).cuda()
for lt in (0, 1, 2, 4, 6):
    for (lbo, sbo) in [(N * 32, 128), (128, N * 32), (16, 1024), (1024, 16), (512, 1024)]:
        g = run(M, N, K, a_img, b_lin, 0, 1, lbo, sbo, 0, lt)[:8]
        print(f"lt={lt} lbo={lbo} sbo={sbo} nonzero={int((g != 0).sum())}",
              "k0:", g[0, :10].astype(int).tolist(), "n0:", g[:, 0].astype(int).tolist(), "n=8,16,32,64:", [int(g[0, n]) for n in (8, 16, 32, 64)])

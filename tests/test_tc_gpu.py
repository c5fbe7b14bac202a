"""Pins the tcgen05 conventions of csrc/tc_tile.cuh: interleaved images as K-major and MN-major operands."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def to_img(mat):
    """[rows, cols] -> interleaved image (flat), off(r,c) = (r/8)*(K*8) + (c/4)*32 + (r%8)*4 + (c%4), K = cols."""
    R, K = mat.shape
    out = np.zeros(R * K, np.float32)
    r, c = np.meshgrid(np.arange(R), np.arange(K), indexing="ij")
    off = (r >> 3) * (K * 8) + (c >> 2) * 32 + (r & 7) * 4 + (c & 3)
    out[off.reshape(-1)] = mat.reshape(-1)
    return out


# K-major operands only: tf32 MN-major operands with the no-swizzle layouts read back as zeros on this part
# (probed with scripts/tc_probe.py), so the kernels write explicit transposed images instead.
@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(64, 128, 160, 0, 0), (128, 64, 32, 0, 0), (64, 256, 128, 0, 0),
                                              (128, 128, 64, 0, 0), (128, 160, 64, 0, 0), (64, 16, 256, 0, 0),
                                              (64, 256, 16, 0, 0)])
def test_tcgen05_tile_gemm(M, N, K, a_mn, b_mn):
    from reinlife_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    a_img = to_img(A.T.copy() if a_mn else A)       # MN-major image: k rows x mn columns
    b_img = to_img(B.T.copy() if b_mn else B)
    ta, tb = torch.from_numpy(a_img).cuda(), torch.from_numpy(b_img).cuda()
    d = torch.zeros((M, N), device="cuda")
    _lib.check(lib.rl_tc_gemm_test(C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(d.data_ptr()),
                                   M, N, K, a_mn, b_mn, None))
    torch.cuda.synchronize()
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = d.cpu().numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err < 2e-3, err      # tf32: 10-bit mantissa products, fp32 accumulation


def test_tensor_core_learn_matches_fp32_learn():
    """rl_brain_learn_tc (tcgen05 tf32) and rl_brain_learn_h (tcgen05 fp16 operands) vs rl_brain_learn (fp32 FMA) on the same events: gradients within 1% of the
    gradient scale, losses / priorities within 2e-2 relative."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_learn_gpu import _mk, _fake_events, _fill_ring
    from brain_golden_util import golden, state_dict
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain, ReplayRings
    z = golden()
    rng = np.random.default_rng(3)
    NW, cap = 7, 256
    per_world = [3, 1, 0, 5, 2, 4, 6]
    vw, rows = _mk(NW)
    w0, tgt = state_dict("train_perd3qn/w0"), state_dict("train_perd3qn/target")
    rp = ReplayRings(NW, cap, "cuda")
    obs_all = z["obs"]
    for w in range(NW):
        n = 200
        o = obs_all[rng.integers(0, 512, n)]; no = obs_all[rng.integers(0, 512, n)]
        a = rng.integers(0, 8, n); r = rng.choice([0.0, 0.2, 0.5, -3.0, -20.0], n); d = (r < 0).astype(np.float64)
        _fill_ring(rp, w, o, a, r, no, d)
    n_ev = _fake_events(vw, rows, per_world)
    sidx = torch.from_numpy(rng.integers(0, 200, size=(n_ev, 64)).astype(np.int32)).cuda()
    out = {}
    for mode in ("fp32", "tf32", "fp16", "fp16p"):       # fp16p = rl_brain_learn_p (two events per CTA iteration)
        brain = DeviceBrain(0, w0, "cuda", lr=1e-3, gamma=0.99)
        brain.use_fp16 = mode in ("fp16", "fp16p")
        brain.load_state_dict(tgt, target=True)
        brain.alloc_learn(rows.row_cap)
        brain.sample_idx[:n_ev] = sidx
        if mode == "fp32":
            _lib.check(vw.lib.rl_brain_learn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                             C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs), vw._stream()))
        elif mode == "tf32":
            brain.build_wimg(vw._stream())
            _lib.check(vw.lib.rl_brain_learn_tc(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                                                C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs),
                                                C.c_void_p(brain.wimg_e.data_ptr()), C.c_void_p(brain.wimg_t.data_ptr()), vw._stream()))
        else:
            brain.build_wimg(vw._stream())
            fn = vw.lib.rl_brain_learn_p if mode == "fp16p" else vw.lib.rl_brain_learn_h
            _lib.check(fn(C.byref(vw.cfg), C.byref(rows.bufs), 0, C.byref(rp.bufs),
                          C.c_void_p(brain.sample_idx.data_ptr()), C.byref(brain.learn_bufs),
                          C.c_void_p(brain.wimg_eh.data_ptr()), C.c_void_p(brain.wimg_th.data_ptr()), vw._stream()))
        torch.cuda.synchronize()
        out[mode] = (brain.grad.cpu().numpy().copy(), brain.loss[:n_ev].cpu().numpy().copy(), brain.new_prio[:n_ev].cpu().numpy().copy())
    g32, l32, p32 = out["fp32"]
    from reinlife_b200.Models import packing
    d = packing.dims(0)
    m = packing.grad_mask(0)
    nt = len(g32) - 4
    for mode in ("tf32", "fp16", "fp16p"):       # all carry 11 significant bits per operand, fp32 accumulation
        gtc, ltc, ptc = out[mode]
        assert gtc[nt] == n_ev
        for name, lo, hi in (("W1", 0, d.off_b1), ("b1", d.off_b1, d.off_w2t), ("W2", d.off_w2t, d.off_b2), ("b2", d.off_b2, d.off_wh),
                             ("Wh", d.off_wh, d.off_bh), ("bh", d.off_bh, d.off_bh + 9)):
            a, b = g32[lo:hi] * m[lo:hi], gtc[lo:hi] * m[lo:hi]
            scale = np.abs(a).max()
            err = np.abs(a - b).max() / scale
            assert err < 1e-2, (mode, name, err, scale)
        np.testing.assert_allclose(ltc, l32, rtol=2e-2, atol=1e-3, err_msg=mode)
        np.testing.assert_allclose(ptc, p32, rtol=2e-2, atol=2e-2, err_msg=mode)


def test_tensor_core_act_matches_fp32_act():
    """rl_brain_act_tc (tcgen05 tf32 forward) and rl_brain_act_h (fp16 operands) vs rl_brain_act_all (fp32 FMA) on real observations with the pretrained
    PERD3QN weights: Q values within 2e-2 of the Q scale, >= 99% identical greedy actions, identical exploration draws
    (epsilon = 0.3: wherever the fp32 path explored, both paths pick the same random action)."""
    from brain_golden_util import golden, state_dict
    from reinlife_b200 import _lib
    from reinlife_b200.brains import DeviceBrain
    from reinlife_b200.World.vecworld import VecWorld
    from reinlife_b200.rows import RowLists
    NW = 24
    vw = VecWorld(NW, 30, 30, 2, max_agents=100, seed=12)
    rows = RowLists(vw)
    vw.enable_obs_fp16()                     # "fp16t": rows gathered by TMA from the float16 copies the World kernels emit
    vw.reset(); vw.top_up(100)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for _ in range(3):                       # a few steps so that observations are not the reset ones
        vw.set_actions(torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g))
        vw.step(); vw.update(); vw.top_up(100)
    rows.build(kinds_mask=1)
    brains = [DeviceBrain(0, state_dict("perd3qn"), "cuda"), DeviceBrain(0, state_dict("d3qn"), "cuda")]
    eps = torch.tensor([0.3, 0.0], dtype=torch.float64, device="cuda")
    descs = (_lib.BrainAct * 2)(*[b.act_desc(_lib.ACT_DUELING, eps.data_ptr() + 8 * i) for i, b in enumerate(brains)])
    out = {}
    import os
    for mode in ("fp32", "tf32", "fp16", "fp16p", "fp16t"):       # fp16p / fp16t = rl_brain_act_p (batch-major 128-row tiles, the default)
        os.environ.pop("RL_ACT_NO_TMA", None)
        if mode == "fp16p":
            os.environ["RL_ACT_NO_TMA"] = "1"                     # rows converted from obs_state by the gather warps
        q = torch.zeros((2, rows.row_cap, 8), device="cuda")
        vw.rec[:, :, 13] = 255
        if mode == "fp32":
            _lib.check(vw.lib.rl_brain_act_all(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), descs, 2, C.c_uint64(7),
                                               C.c_void_p(q.data_ptr()), None, vw._stream()))
        else:
            for i, b in enumerate(brains):
                b.use_fp16 = mode != "tf32"
                b.build_wimg(vw._stream())
                if mode == "tf32":
                    _lib.check(vw.lib.rl_brain_act_tc(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), i, C.byref(descs[i]),
                                                      C.c_void_p(b.wimg_e.data_ptr()), C.c_uint64(7), C.c_void_p(q.data_ptr()), vw._stream()))
                else:
                    fn = vw.lib.rl_brain_act_p if mode in ("fp16p", "fp16t") else vw.lib.rl_brain_act_h
                    _lib.check(fn(C.byref(vw.cfg), C.byref(vw.bufs), C.byref(rows.bufs), i, C.byref(descs[i]),
                                  C.c_void_p(b.wimg_eh.data_ptr()), C.c_uint64(7), C.c_void_p(q.data_ptr()), vw._stream()))
        torch.cuda.synchronize()
        out[mode] = (q.cpu().numpy(), vw.rec[:, :, 13].cpu().numpy().view(np.int8).copy())
    os.environ.pop("RL_ACT_NO_TMA", None)
    a32 = out["fp32"][1]
    listed = a32 != -1
    for i in range(2):                         # same float16 operands either way: bit-identical Q values and actions
        n = int(rows.total[i * 3])
        assert np.array_equal(out["fp16p"][0][i, :n], out["fp16t"][0][i, :n])
    assert np.array_equal(out["fp16p"][1], out["fp16t"][1])
    for mode in ("tf32", "fp16", "fp16p", "fp16t"):
        n_tot = 0
        for i in range(2):
            n = int(rows.total[i * 3])
            n_tot += n
            q32, qtc = out["fp32"][0][i, :n], out[mode][0][i, :n]
            assert n > 500 and np.abs(q32).max() > 1.0
            assert np.abs(q32 - qtc).max() < 2e-2 * np.abs(q32).max(), mode
            assert (q32.argmax(1) == qtc.argmax(1)).mean() >= 0.99, mode
        atc = out[mode][1]
        assert listed.sum() == n_tot and ((atc != -1) == listed).all(), mode
        assert (a32[listed] == atc[listed]).mean() >= 0.99, mode


def _himg16(mat):
    """fp16 [R][W] -> interleaved no-swizzle image (8 rows x 16 B core matrices), csrc/tc_bm.cuh::himg."""
    R, W = mat.shape
    out = np.zeros(R * W, np.float16)
    r, c = np.meshgrid(np.arange(R), np.arange(W), indexing="ij")
    out[((r >> 3) * (W * 8) + (c >> 3) * 64 + (r & 7) * 8 + (c & 7)).reshape(-1)] = mat.reshape(-1)
    return out


def _sw128(mat):
    """fp16 [R][W] -> ceil(W / 64) K blocks of [R][128 B], 16-byte unit index ^= (row & 7): csrc/tc_bm.cuh::ximg, what a TMA
    tile::gather4 with CU_TENSOR_MAP_SWIZZLE_128B writes."""
    R, W = mat.shape
    out = np.zeros(((W + 63) // 64) * R * 64, np.float16)
    r, c = np.meshgrid(np.arange(R), np.arange(W), indexing="ij")
    out[((c // 64) * (R * 64) + r * 64 + ((((c % 64) // 8) ^ (r & 7)) * 8) + (c % 8)).reshape(-1)] = mat.reshape(-1)
    return out


def _gemm_hx(a_img, b_img, M, N, K, ageo, bgeo, a_mn, b_mn, akblk, bkblk, alay, blay):
    from reinlife_b200 import _lib
    lib = _lib.load()
    ta, tb = torch.from_numpy(a_img).cuda(), torch.from_numpy(b_img).cuda()
    d = torch.zeros((M, N), device="cuda")
    _lib.check(lib.rl_tc_gemm_test_hx(C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(d.data_ptr()), M, N, K,
                                      a_img.size, b_img.size, *ageo, *bgeo, a_mn, b_mn, akblk, bkblk, alay, blay, None))
    torch.cuda.synchronize()
    return d.cpu().numpy()


def test_swizzle128_operand_conventions_and_tma_gather4():
    """What k_learn_dueling_p / k_act_dueling_p rely on for their gathered input rows (csrc/tc_bm.cuh):
    (1) a SWIZZLE_128B image [128 rows][160 halves] (three K blocks) is a valid K-major A operand (L1: X W1^T, W1 no-swizzle);
    (2) the SAME image is a valid MN-major B operand with N = image columns, K = image rows (dW1 = dH1^T X, N = 160 / 80 / 64);
    (3) cp.async.bulk.tensor tile::gather4 of 128 ring rows by index (box {64 columns, 1 row}, SWIZZLE_128B) lands exactly that
        image, columns >= 160 as zeros."""
    from reinlife_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(3)
    M, N, K = 128, 128, 160
    X = rng.standard_normal((M, K)).astype(np.float16); W = rng.standard_normal((N, K)).astype(np.float16)
    got = _gemm_hx(_sw128(X), _himg16(W), M, N, K, (16, 1024, 32), (128, K * 16, 256), 0, 0, M * 128, 0, 2, 0)
    want = X.astype(np.float64) @ W.astype(np.float64).T
    assert np.abs(got - want).max() < 1e-5 * np.abs(want).max()
    dH1 = rng.standard_normal((128, 128)).astype(np.float16)
    want = dH1.astype(np.float64).T @ X.astype(np.float64)
    for Nx in (160, 80, 64, 128):
        got = _gemm_hx(_himg16(dH1), _sw128(X), 128, Nx, 128, (128 * 16, 128, 2 * 128 * 16), (16384, 1024, 2048), 1, 1, 0, 0, 0, 2)
        assert np.abs(got - want[:, :Nx]).max() < 1e-5 * np.abs(want).max(), Nx
    n_rows = 5000
    ring = rng.standard_normal((n_rows, 160)).astype(np.float16)
    idx = rng.integers(0, n_rows, 128).astype(np.int32)
    idx[:4] = (0, n_rows - 1, 17, 17)                                       # first / last row, a duplicate
    tr, ti = torch.from_numpy(ring).cuda(), torch.from_numpy(idx).cuda()
    out = torch.zeros(3 * 16384 // 2, dtype=torch.float16, device="cuda")
    _lib.check(lib.rl_tma_gather_test(C.c_void_p(tr.data_ptr()), C.c_longlong(n_rows), C.c_void_p(ti.data_ptr()), C.c_void_p(out.data_ptr()), 1))
    want = _sw128(np.concatenate([ring[idx], np.zeros((128, 32), np.float16)], axis=1))
    assert np.array_equal(out.cpu().numpy().view(np.uint16), want.view(np.uint16))

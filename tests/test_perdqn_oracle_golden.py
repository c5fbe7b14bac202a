"""Pins the PERDQN part of oracle/brain_oracle.py against tests/golden/brain_golden3.npz: a 150-store / 3-train /
250-store (ring wrap) / 1-train run of the reference's PERDQNAgent (Models/PERDQN.py) recorded call by call."""
import numpy as np

from oracle import brain_oracle as bo
from perdqn_golden_util import golden3, meta3, sd3, transitions


def test_perdqn_forward_matches_reference_module():
    z = golden3()
    np.testing.assert_allclose(bo.perdqn_forward(sd3("fwd/w"), z["fwd/obs"]), z["fwd/q"], rtol=1e-4, atol=1e-4)


def test_perdqn_run_replays_bit_exact_tree_and_samples():
    z, M = golden3(), meta3()
    cap = M["capacity"]
    w = {k: v.copy() for k, v in sd3("run/w0").items()}
    target = sd3("run/target")
    m = {k: np.zeros_like(v) for k, v in w.items()}
    v2 = {k: np.zeros_like(v) for k, v in w.items()}
    tree = bo.SumTreeOracle(cap)
    data = [None] * cap
    eps, n_add, n_upd, n_train = 1.0, 0, 0, 0

    def store(lo, hi):
        nonlocal n_add
        s, a, r, s2, d = transitions(lo, hi)
        err = bo.perdqn_store_error(w, target, s, a, r, s2, d, gamma=M["gamma"])
        assert np.array_equal(err, z["run/add_err"][lo:hi]) and not err.any()         # the aliasing quirk: always 0
        # the priority rule on the reference's own error: torch float32 pow vs numpy float32 pow, <= 1 ulp apart
        leaf = z["run/add_leaf"][lo:hi]
        np.testing.assert_allclose(bo.perdqn_priority(z["run/add_err"][lo:hi]).astype(np.float64), leaf, rtol=2e-7)
        for i in range(lo, hi):                       # tree arithmetic: replayed with the reference's leaves -> bit-exact
            data[tree.add(leaf[i - lo])] = i
            n_add += 1

    def train():
        nonlocal eps, n_upd, n_train
        if eps > M["eps_min"]:
            eps -= M["eps_decay"]
        us = iter(z[f"run/sample{n_train}/u"])
        slots, idxs, isw = tree.sample(64, lambda: float(next(us)))
        assert next(us, None) is None, "the oracle consumed fewer uniform draws than the reference"
        assert idxs == list(z[f"run/sample{n_train}/idx"])
        np.testing.assert_allclose(isw, z[f"run/sample{n_train}/isw"], rtol=1e-12)
        ids = [data[s_] for s_ in slots]
        s, a, r, s2, d = (x[ids] for x in transitions(0, len(z["run/action"])))
        grads, _, errors = bo.perdqn_event_grads(w, target, s, a, r, s2, d, isw, gamma=M["gamma"])
        ref_err = z["run/upd_err"][n_upd:n_upd + 64]
        np.testing.assert_allclose(errors, ref_err, rtol=1e-4, atol=2e-5)
        assert list(z["run/upd_idx"][n_upd:n_upd + 64]) == idxs
        for idx, e in zip(idxs, ref_err):             # Memory.update in batch order, duplicates included
            tree.update(idx, np.float64(bo.perdqn_priority(e)))
        n_upd += 64
        bo.adam_step(w, grads, m, v2, n_train + 1, lr=M["lr"])
        ref_w = sd3(f"run/step{n_train}")
        for k in w:
            np.testing.assert_allclose(w[k], ref_w[k], rtol=0, atol=3e-6, err_msg=f"train {n_train} {k}")
        n_train += 1

    def check(name):
        assert np.array_equal(tree.tree, z[f"run/{name}/tree"]), name          # float64 bit patterns
        write, n_entries, beta, ref_eps = z[f"run/{name}/scal"]
        assert (tree.write, tree.n_entries) == (int(write), int(n_entries))
        assert tree.beta == beta and abs(eps - ref_eps) < 1e-15

    store(0, 150); check("p0_store150")
    for k in range(3):
        train(); check(f"p{k + 1}_train")
    store(150, 400); check("p4_store250")
    train(); check("p5_train")
    assert tree.n_entries == cap and tree.write == 400 % cap


def test_sumtree_redraws_unfilled_leaves():
    """Memory.sample redraws while the leaf holds no data (PERDQN.py:291-295): with a single stored item every stratum
    ends on it, whatever the draw."""
    t = bo.SumTreeOracle(7)
    t.add(0.5)
    us = iter(np.linspace(0.0, 1.0, 64, endpoint=False))
    slots, idxs, w = t.sample(8, lambda: float(next(us)))
    assert slots == [0] * 8 and idxs == [6] * 8 and np.all(w == 1.0)


def test_perdqn_rule():
    q = [0.1, 0.9, 0.9, 0.2]
    assert bo.perdqn_rule(q, 0.5, 0.5, 7) == 7 and bo.perdqn_rule(q, 0.5, 0.5000001, 7) == 1


def test_host_brain_draws_the_reference_initial_weights():
    """reinlife_b200.Models.PERDQN() consumes torch's global RNG exactly like the reference constructor (model with
    xavier weights, then the throw-away target_model draw) and round-trips through the padded kernel layout."""
    import torch
    from reinlife_b200.Models import PERDQN, packing
    z = golden3()
    torch.manual_seed(123)
    b = PERDQN()
    want = sd3("init/w")
    assert b.method == "PERDQN" and abs(b.epsilon_decay - 0.99 / 5000) < 1e-18 and b.train_start == 1000
    for k, v in b.model.state_dict().items():
        assert np.array_equal(v.numpy(), want[k]), k
    assert np.array_equal(torch.rand(4).numpy(), z["init/next_rand"])
    flat = packing.pack(packing.PERDQN, want)
    back = packing.unpack(packing.PERDQN, flat)
    assert list(back) == list(want) and all(np.array_equal(back[k].numpy(), want[k]) for k in want)
    m, d = packing.grad_mask(packing.PERDQN), packing.dims(packing.PERDQN)
    assert int(m.sum()) == 153 * 64 + 64 + 64 * 64 + 64 + 64 * 8 + 8 == 14536          # 14 536 parameters
    assert not flat[:d.n_train][m == 0].any()

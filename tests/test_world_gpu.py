"""GPU parity of the World kernels (through the C ABI) against (a) the golden vectors minted from the
unmodified reference and (b) the C oracle on seeded batches of worlds.  Bit-exact: cell types, agent
records, float32(reward), float32(observation)."""
import numpy as np
import pytest
import torch

from golden_util import load_cases, rec_equal, rec_diff

pytestmark = pytest.mark.gpu


def _vw(*a, **k):
    from reinlife_b200.World.vecworld import VecWorld
    return VecWorld(*a, **k)


def _cmp_world(tag, vw, w, typ, rec, reward=None, obs=None, which="state"):
    n = len(rec)
    assert int(vw.n_agents[w]) == n, (tag, int(vw.n_agents[w]), n)
    assert (vw.type[w].cpu().numpy() == np.asarray(typ).reshape(-1)).all(), tag
    got = vw.rec_host()[w, :n]
    assert rec_equal(got, rec), (tag, rec_diff(got, rec))
    if reward is not None:
        g = vw.reward[w, :n].cpu().numpy()
        assert (g.view(np.uint32) == reward.astype(np.float32).view(np.uint32)).all(), (tag, g, reward)
    if obs is not None:
        t = vw.obs_prime if which == "prime" else vw.obs_state
        g = t[w, :n].cpu().numpy()
        assert (g[:, :153].view(np.uint32) == obs.astype(np.float32).view(np.uint32)).all(), (tag, which)
        assert (g[:, 153:] == 0).all(), tag


@pytest.mark.parametrize("phase", ["reset", "step", "update", "topup"])
def test_world_kernels_match_reference_golden(phase):
    n = 0
    for k, c in enumerate(load_cases()):
        if c["phase"] != phase:
            continue
        vw = _vw(1, c["height"], c["width"], c["n_genes"], c["max_agents"], seed=c["seed"], world_id0=c["world"],
                 limit_reproduction=c["limit_reproduction"], incentivize_killing=c["incentivize_killing"])
        if phase == "reset":
            vw.reset()
        else:
            vw.load_host(0, c["in_type"], c["in_rec"])
            vw.t = c["t"]
            if phase == "step":
                vw.t = c["t"] - 1
                vw.step()
            elif phase == "update":
                vw.update()
            else:
                vw.top_up(c["target"])
        torch.cuda.synchronize()
        _cmp_world((k, phase), vw, 0, c["out_type"], c["out_rec"],
                   reward=c["out_reward"] if phase == "step" else None, obs=c["out_obs"],
                   which="prime" if phase == "step" else "state")
        n += 1
    assert n > 0


@pytest.mark.parametrize("H,W,G,NW,target,steps", [(30, 30, 2, 64, 100, 12), (9, 7, 3, 96, 30, 25),
                                                   (60, 60, 2, 8, 400, 6), (30, 30, 5, 32, 0, 40)])
def test_world_kernels_match_oracle_trajectories(H, W, G, NW, target, steps):
    from oracle.world_oracle import OracleWorlds
    seed = 1234 + H
    vw = _vw(NW, H, W, G, max_agents=max(target, 20), seed=seed, world_id0=17)
    ow = OracleWorlds(NW, H, W, G, max_agents=max(target, 20), seed=seed, world_id0=17)
    vw.reset(); ow.reset()
    rng = np.random.default_rng(seed)

    def compare(tag, which, reward=False):
        torch.cuda.synchronize()
        for w in range(NW):
            n = int(ow.n[w])
            _cmp_world((tag, w), vw, w, ow.type[w], ow.rec[w, :n], reward=ow.reward[w, :n] if reward else None,
                       obs=ow.obs[w, :n], which=which)

    compare("reset", "state")
    if target:
        vw.top_up(target); ow.top_up(target)
        compare("topup0", "state")
    for s in range(steps):
        acts = rng.integers(0, 8, size=(NW, ow.S)).astype(np.int8)
        ow.set_actions(acts); vw.set_actions(acts)
        ow.step(); vw.step()
        compare(("step", s), "prime", reward=True)
        ow.update(); vw.update()
        compare(("update", s), "state")
        if target:
            vw.top_up(target); ow.top_up(target)
            compare(("topup", s), "state")
    assert int(vw.status.max()) == 0


def test_world_sharding_is_world_id_pure():
    """Worlds 8..15 of a 16-world batch == an 8-world shard with world_id0=8 (multi-GPU sharding rule)."""
    full = _vw(16, 12, 12, 2, 30, seed=5)
    part = _vw(8, 12, 12, 2, 30, seed=5, world_id0=8)
    full.reset(); part.reset()
    full.top_up(30); part.top_up(30)
    rng = np.random.default_rng(0)
    for s in range(10):
        acts = rng.integers(0, 8, size=(16, full.S)).astype(np.int8)
        full.set_actions(acts); part.set_actions(acts[8:])
        full.step(); part.step(); full.update(); part.update()
    torch.cuda.synchronize()
    assert torch.equal(full.type[8:], part.type)
    assert torch.equal(full.n_agents[8:], part.n_agents)
    assert torch.equal(full.obs_state[8:], part.obs_state)


def test_full_size_world_batch_properties_and_oracle_spot_checks():
    """BASELINE.json configs[2] size (4096 worlds, 30x30, saturated to 100 agents): size-independent invariants on every
    world + bit-exact oracle parity on three 16-world slices of the batch (first, middle, last) + run-to-run determinism."""
    from oracle.world_oracle import OracleWorlds
    NW, H, W, G, target, steps, seed = 4096, 30, 30, 2, 100, 4, 77
    slices = [0, 2040, 4080]

    def run():
        vw = _vw(NW, H, W, G, max_agents=target, seed=seed)
        ows = [OracleWorlds(16, H, W, G, max_agents=target, seed=seed, world_id0=s0) for s0 in slices]
        vw.reset(); vw.top_up(target)
        for ow in ows:
            ow.reset(); ow.top_up(target)
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        sums = []
        for s in range(steps):
            acts = torch.randint(0, 8, (NW, vw.S), device="cuda", dtype=torch.int8, generator=g)
            vw.set_actions(acts)
            ah = acts.cpu().numpy()
            vw.step()
            for ow, s0 in zip(ows, slices):
                ow.set_actions(ah[s0:s0 + 16]); ow.step()
            torch.cuda.synchronize()
            for ow, s0 in zip(ows, slices):
                for w in range(16):
                    n = int(ow.n[w])
                    _cmp_world(("full step", s, s0 + w), vw, s0 + w, ow.type[w], ow.rec[w, :n], reward=ow.reward[w, :n],
                               obs=ow.obs[w, :n], which="prime")
            vw.update(); vw.top_up(target)
            for ow in ows:
                ow.update(); ow.top_up(target)
            torch.cuda.synchronize()
            for ow, s0 in zip(ows, slices):
                for w in range(16):
                    n = int(ow.n[w])
                    _cmp_world(("full update", s, s0 + w), vw, s0 + w, ow.type[w], ow.rec[w, :n], obs=ow.obs[w, :n], which="state")
            # ---- invariants on ALL worlds
            n = vw.n_agents.long()
            is_agent = vw.type == 3
            assert torch.equal(is_agent.sum(1), n)                                   # one list entry per AGENT cell
            rec = vw.rec.view(NW, vw.S, 16)
            cell = rec[:, :, 0].long() | (rec[:, :, 1].long() << 8)
            slot = torch.arange(vw.S, device="cuda")[None, :]
            live = slot < n[:, None]
            nxt = torch.roll(cell, -1, 1)
            assert bool(((cell < nxt) | ~(slot + 1 < n[:, None])).all())             # row-major (strictly increasing cells)
            assert bool((torch.gather(vw.type.long(), 1, cell.clamp(max=H * W - 1))[live] == 3).all())
            assert int(n.min()) == target                                            # saturated by the top-up
            assert bool((vw.obs_state[:, :, 153:] == 0).all()) and bool(torch.isfinite(vw.reward).all())
            o = vw.obs_state[live]
            assert bool(((o[:, :49] == 0) | (o[:, :49] == 0.5) | (o[:, :49] == 1) | (o[:, :49] == -1)).all())   # food plane codes
            assert bool(((o[:, 98:147] == 0) | (o[:, 98:147] == 1) | (o[:, 98:147] == -1)).all())              # gene plane codes
            sums.append((int(vw.type.long().sum()), int(cell[live].sum()), float(vw.obs_state.double().sum()), float(vw.reward.double().sum())))
        assert int(vw.status.max()) == 0
        return sums

    a = run()
    b = run()
    assert a == b                                                                    # bit-reproducible run to run


def test_fused_update_top_up_equals_update_then_top_up():
    """rl_world_update_top_up (one launch: the benchmark loops' update_env + saturate) leaves exactly the state that
    rl_world_update followed by rl_world_top_up leaves: cell types, agent records, counts, observations -- bit for bit,
    over a run with deaths, births and a moving saturation target."""
    from reinlife_b200.World.vecworld import VecWorld
    for (H, W, G, NW, target) in ((30, 30, 2, 40, 100), (9, 7, 3, 24, 20), (12, 17, 5, 16, 60)):
        a = VecWorld(NW, H, W, G, max_agents=target, seed=21)
        b = VecWorld(NW, H, W, G, max_agents=target, seed=21)
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        for vw in (a, b):
            vw.reset(); vw.top_up(target)
        for t in range(25):
            act = torch.randint(0, 8, (NW, a.S), device="cuda", dtype=torch.int8, generator=g)
            tgt = target if t % 3 else max(1, target // 2)          # below the current count on some steps: no placement
            for vw in (a, b):
                vw.set_actions(act); vw.step()
            a.update(); a.top_up(tgt)
            b.update(top_up=tgt)
            torch.cuda.synchronize()
            assert torch.equal(a.type, b.type) and torch.equal(a.n_agents, b.n_agents), (H, W, t)
            n = a.n_agents.cpu().numpy()
            ra, rb = a.rec.cpu().numpy(), b.rec.cpu().numpy()
            oa, ob = a.obs_state.view(torch.int32).cpu().numpy(), b.obs_state.view(torch.int32).cpu().numpy()
            for w in range(NW):
                assert (ra[w, :n[w]] == rb[w, :n[w]]).all(), (H, W, t, w)
                assert (oa[w, :n[w]] == ob[w, :n[w]]).all(), (H, W, t, w)


def test_float16_observation_copies():
    """enable_obs_fp16(): reset / step / update / top-up also write float16 copies of the rows they write in float32 -- exactly
    float16(row) with element 159 = 1.0 for every listed agent -- and change nothing else (same float32 rows, records, cells)."""
    from reinlife_b200.World.vecworld import VecWorld
    NW, H, W, G, target = 24, 30, 30, 2, 100
    a = VecWorld(NW, H, W, G, max_agents=target, seed=4)
    b = VecWorld(NW, H, W, G, max_agents=target, seed=4)
    b.enable_obs_fp16()
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    for vw in (a, b):
        vw.reset(); vw.top_up(target)

    def check(which):
        torch.cuda.synchronize()
        assert torch.equal(a.type, b.type) and torch.equal(a.n_agents, b.n_agents) and torch.equal(a.rec, b.rec)
        f32, f16 = (b.obs_prime, b.obs_prime_h) if which else (b.obs_state, b.obs_state_h)
        assert torch.equal(f32, a.obs_prime if which else a.obs_state)
        listed = torch.arange(b.S, device="cuda")[None, :] < b.n_agents[:, None]
        want = f32.half().clone(); want[..., 159] = 1.0
        assert torch.equal(f16[listed], want[listed]), which

    check(0)
    for t in range(8):
        act = torch.randint(0, 8, (NW, a.S), device="cuda", dtype=torch.int8, generator=g)
        for vw in (a, b):
            vw.set_actions(act); vw.step()
        check(1)
        a.update(); b.update(); check(0)
        a.top_up(target); b.top_up(target); check(0)
        for vw in (a, b):
            vw.set_actions(act); vw.step()
        a.update(top_up=target); b.update(top_up=target); check(0)

"""Checkpoint layout of reinlife_b200.Saver == the reference's (Helpers/saver.py:58-194, World/entities.py:224-242):
directory tree, file names, state_dict key names / shapes per method, parameters json.  Host-only (no GPU): brains
that are not bound to an Environment keep their networks on the host."""
import json
import os
from datetime import date

import numpy as np
import torch

from brain_golden_util import state_dict
from perdqn_golden_util import sd3


def _brains():
    from reinlife_b200.Models import D3QN, DQN, PERD3QN, PERDQN, PPO
    torch.manual_seed(0)
    return [PERD3QN(), D3QN(), DQN(max_epi=10), PPO(), PERDQN()]


def test_saver_writes_the_reference_layout(tmp_path, monkeypatch):
    import reinlife_b200 as rl
    from reinlife_b200.Helpers.saver import Saver, _SavedAgent
    assert rl.Saver is Saver
    monkeypatch.chdir(tmp_path)
    brains = _brains()
    agents = [_SavedAgent(g, b) for g, b in enumerate(brains)]
    results = {"Avg Population Size": {"0": [1.0, 2.0]}}
    settings = {"Update interval": 500, "Width": 30, "Height": 30, "Max agents": 100, "Families": True}
    written = Saver("experiments").save(agents, True, results, settings, None)
    exp = tmp_path / "experiments" / f"{date.today()}_V1"
    assert sorted(p.name for p in exp.iterdir()) == ["D3QN", "DQN", "PERD3QN", "PERDQN", "PPO", "results.json", "settings.json"]
    assert json.load(open(exp / "results.json")) == results and json.load(open(exp / "settings.json")) == settings
    ref_keys = {"PERD3QN": state_dict("perd3qn"), "D3QN": state_dict("d3qn"), "DQN": state_dict("dqn"),
                "PPO": state_dict("ppo"), "PERDQN": sd3("fwd/w")}
    for g, b in enumerate(brains):
        f = exp / b.method / f"brain_gene_{g}.pt"
        assert str(f) in written and f.exists()
        sd = torch.load(f, map_location="cpu")
        want = ref_keys[b.method]
        assert list(sd.keys()) == list(want.keys())                       # the reference's names, in module order
        assert all(tuple(sd[k].shape) == want[k].shape and sd[k].dtype == torch.float32 for k in want)
        params = json.load(open(exp / b.method / f"parameters_gene_{g}.json"))
        assert params["method"] == params["_method"] == b.method and params["input_dim"] == 153 and params["one"] == 1
        assert not any(k.isupper() for k in params) and "training" in params
    assert json.load(open(exp / "PERDQN" / "parameters_gene_4.json"))["train_start"] == 1000
    # a second experiment on the same day -> _V2 (saver.py:121-127); the checkpoint loads back through load_model
    Saver("experiments").save(agents[:1], True, results, settings, None)
    assert (tmp_path / "experiments" / f"{date.today()}_V2" / "PERD3QN" / "brain_gene_0.pt").exists()
    from reinlife_b200.Models import PERDQN
    again = PERDQN(load_model=str(exp / "PERDQN" / "brain_gene_4.pt"), training=False)
    for k, v in brains[4].model.state_dict().items():
        assert np.array_equal(again.model.state_dict()[k].numpy(), v.numpy())


def test_saver_numbers_brains_per_method_without_families(tmp_path, monkeypatch):
    from reinlife_b200.Helpers.saver import Saver, _SavedAgent
    from reinlife_b200.Models import PPO, PERD3QN
    monkeypatch.chdir(tmp_path)
    agents = [_SavedAgent(7, PPO()), _SavedAgent(9, PERD3QN()), _SavedAgent(12, PPO())]
    Saver("experiments").save(agents, False, {}, {}, None)
    exp = tmp_path / "experiments" / f"{date.today()}_V1"
    assert sorted(os.listdir(exp / "PPO")) == ["brain_1.pt", "brain_2.pt", "parameters_1.json", "parameters_2.json"]
    assert sorted(os.listdir(exp / "PERD3QN")) == ["brain_1.pt", "parameters_1.json"]

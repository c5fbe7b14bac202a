"""state_dict <-> kernel-layout parameter buffers (layout contract: include/reinlife_b200.h, "Brains").

Key names and shapes are the reference's (Models/PERD3QN.py:185-196, Models/D3QN.py:148-159,
Models/DQN.py:119-124, Models/PPO.py:96-99) so that `pretrained/**/brain_gene_*.pt` load unchanged
and checkpoints written here load in the reference.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from .. import _lib

DUELING, DQN, PPO = 0, 1, 2       # = rl_model_kind
PERDQN = 3                        # host-side kind: Models/PERDQN.py:311-323 (153-64-64-8) carried in the DQN kernel layout,
                                  # first hidden layer zero-padded from 64 to 128 units (include/reinlife_b200.h, "PERDQN")
K1 = 160


def device_kind(kind):
    """rl_model_kind the kernels see for a host-side kind."""
    return DQN if kind == PERDQN else kind


class ModelDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n1", "n2", "nh", "off_b1", "off_w2t", "off_b2", "off_wh", "off_bh",
                                         "off_w2", "n_train", "n_total")]


_dims_cache = {}


def dims(kind):
    kind = device_kind(kind)
    if kind not in _dims_cache:
        d = ModelDims()
        _lib.check(_lib.load().rl_model_get_dims(C.c_int32(kind), C.byref(d)))
        _dims_cache[kind] = d
    return _dims_cache[kind]


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32) if hasattr(t, "detach") else np.asarray(t, np.float32)


def _layers(kind, sd):
    """-> (W1 [N1,153], b1, W2 [N2,N1], b2, Wh_dense [NH,N2], bh) in nn.Linear (output-major) orientation."""
    if kind == DUELING:
        w1, b1 = _np(sd["fc.weight"]), _np(sd["fc.bias"])
        w2 = np.concatenate([_np(sd["adv_fc1.weight"]), _np(sd["value_fc1.weight"])], 0)
        b2 = np.concatenate([_np(sd["adv_fc1.bias"]), _np(sd["value_fc1.bias"])], 0)
        wh = np.zeros((9, 256), np.float32)
        wh[:8, :128] = _np(sd["adv_fc2.weight"])
        wh[8, 128:] = _np(sd["value_fc2.weight"])[0]
        bh = np.concatenate([_np(sd["adv_fc2.bias"]), _np(sd["value_fc2.bias"])], 0)
    elif kind == DQN:
        w1, b1 = _np(sd["fc1.weight"]), _np(sd["fc1.bias"])
        w2, b2 = _np(sd["fc2.weight"]), _np(sd["fc2.bias"])
        wh, bh = _np(sd["fc3.weight"]), _np(sd["fc3.bias"])
    elif kind == PERDQN:
        w1, b1 = np.zeros((128, 153), np.float32), np.zeros(128, np.float32)
        w1[:64], b1[:64] = _np(sd["fc.0.weight"]), _np(sd["fc.0.bias"])
        w2 = np.zeros((64, 128), np.float32)
        w2[:, :64], b2 = _np(sd["fc.2.weight"]), _np(sd["fc.2.bias"])
        wh, bh = _np(sd["fc.4.weight"]), _np(sd["fc.4.bias"])
    elif kind == PPO:
        w1, b1 = _np(sd["fc1.weight"]), _np(sd["fc1.bias"])
        w2, b2 = _np(sd["fc2.weight"]), _np(sd["fc2.bias"])
        wh = np.concatenate([_np(sd["fc_pi.weight"]), _np(sd["fc_v.weight"])], 0)
        bh = np.concatenate([_np(sd["fc_pi.bias"]), _np(sd["fc_v.bias"])], 0)
    else:
        raise ValueError(kind)
    return w1, b1, w2, b2, wh, bh


def pack(kind, sd):
    d = dims(kind)
    w1, b1, w2, b2, wh, bh = _layers(kind, sd)
    flat = np.zeros(d.n_total, np.float32)
    w1t = np.zeros((K1, d.n1), np.float32)
    w1t[:153] = w1.T
    flat[0:d.off_b1] = w1t.reshape(-1)
    flat[d.off_b1:d.off_b1 + d.n1] = b1
    flat[d.off_w2t:d.off_b2] = w2.T.reshape(-1)
    flat[d.off_b2:d.off_b2 + d.n2] = b2
    flat[d.off_wh:d.off_bh] = wh.T.reshape(-1)
    flat[d.off_bh:d.off_bh + d.nh] = bh
    flat[d.off_w2:d.off_w2 + d.n1 * d.n2] = w2.reshape(-1)
    return flat


def unpack(kind, flat):
    import torch
    d = dims(kind)
    flat = np.asarray(flat, np.float32)
    w1 = flat[0:d.off_b1].reshape(K1, d.n1)[:153].T.copy()
    b1 = flat[d.off_b1:d.off_b1 + d.n1].copy()
    w2 = flat[d.off_w2t:d.off_b2].reshape(d.n1, d.n2).T.copy()
    b2 = flat[d.off_b2:d.off_b2 + d.n2].copy()
    wh = flat[d.off_wh:d.off_bh].reshape(d.n2, d.nh).T.copy()
    bh = flat[d.off_bh:d.off_bh + d.nh].copy()
    t = torch.from_numpy
    if kind == DUELING:
        return OrderedDict([("fc.weight", t(w1)), ("fc.bias", t(b1)),
                            ("adv_fc1.weight", t(w2[:128].copy())), ("adv_fc1.bias", t(b2[:128].copy())),
                            ("adv_fc2.weight", t(wh[:8, :128].copy())), ("adv_fc2.bias", t(bh[:8].copy())),
                            ("value_fc1.weight", t(w2[128:].copy())), ("value_fc1.bias", t(b2[128:].copy())),
                            ("value_fc2.weight", t(wh[8:9, 128:].copy())), ("value_fc2.bias", t(bh[8:9].copy()))])
    if kind == DQN:
        return OrderedDict([("fc1.weight", t(w1)), ("fc1.bias", t(b1)), ("fc2.weight", t(w2)), ("fc2.bias", t(b2)),
                            ("fc3.weight", t(wh)), ("fc3.bias", t(bh))])
    if kind == PERDQN:
        return OrderedDict([("fc.0.weight", t(w1[:64].copy())), ("fc.0.bias", t(b1[:64].copy())),
                            ("fc.2.weight", t(w2[:, :64].copy())), ("fc.2.bias", t(b2)),
                            ("fc.4.weight", t(wh)), ("fc.4.bias", t(bh))])
    return OrderedDict([("fc1.weight", t(w1)), ("fc1.bias", t(b1)), ("fc2.weight", t(w2)), ("fc2.bias", t(b2)),
                        ("fc_pi.weight", t(wh[:8].copy())), ("fc_pi.bias", t(bh[:8].copy())),
                        ("fc_v.weight", t(wh[8:9].copy())), ("fc_v.bias", t(bh[8:9].copy()))])


def grad_mask(kind):
    """1 for trainable entries of the first n_train floats, 0 for padding / structural zeros."""
    d = dims(kind)
    m = np.zeros(d.n_train, np.float32)
    w1t = np.zeros((K1, d.n1), np.float32)
    w1t[:153] = 1
    m[0:d.off_b1] = w1t.reshape(-1)
    m[d.off_b1:d.off_bh + d.nh] = 1
    if kind == DUELING:
        wh = np.zeros((d.n2, d.nh), np.float32)
        wh[:128, :8] = 1
        wh[128:, 8] = 1
        m[d.off_wh:d.off_bh] = wh.reshape(-1)
    if kind == PERDQN:                          # the padded hidden units 64..127 are not parameters
        w1t[:, 64:] = 0
        m[0:d.off_b1] = w1t.reshape(-1)
        m[d.off_b1 + 64:d.off_b1 + 128] = 0
        w2t = np.zeros((d.n1, d.n2), np.float32)
        w2t[:64] = 1
        m[d.off_w2t:d.off_b2] = w2t.reshape(-1)
    return m


def default_init(kind):
    """Initial weights exactly as the reference draws them: torch.nn.Linear modules constructed in the reference's
    attribute order (PERD3QN.py:189-196, DQN.py:122-124, PPO.py:96-99) from torch's global RNG -- so the same
    torch.manual_seed gives the same initial network as the reference class."""
    import torch.nn as nn
    if kind == PERDQN:
        # PERDQN.py:73-74,92-95: DQN(...) = Sequential(Linear, ReLU, Linear, ReLU, Linear) with default init, then
        # model.apply(weights_init): xavier_uniform on every Linear weight in module order (biases keep the default draw)
        layers = [nn.Linear(153, 64), nn.Linear(64, 64), nn.Linear(64, 8)]
        for layer in layers:
            nn.init.xavier_uniform_(layer.weight)
        sd = OrderedDict()
        for i, layer in zip((0, 2, 4), layers):
            sd[f"fc.{i}.weight"] = layer.weight.detach().clone()
            sd[f"fc.{i}.bias"] = layer.bias.detach().clone()
        return sd
    spec = {DUELING: (("fc", 153, 128), ("adv_fc1", 128, 128), ("adv_fc2", 128, 8), ("value_fc1", 128, 128), ("value_fc2", 128, 1)),
            DQN: (("fc1", 153, 128), ("fc2", 128, 64), ("fc3", 64, 8)),
            PPO: (("fc1", 153, 256), ("fc2", 256, 256), ("fc_pi", 256, 8), ("fc_v", 256, 1))}[kind]
    sd = OrderedDict()
    for name, i, o in spec:
        layer = nn.Linear(i, o)
        sd[name + ".weight"] = layer.weight.detach().clone()
        sd[name + ".bias"] = layer.bias.detach().clone()
    return sd

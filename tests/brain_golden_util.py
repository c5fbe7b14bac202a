import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "brain_golden.npz")
_z = {}


def golden():
    if "z" not in _z:
        _z["z"] = np.load(PATH)
    return _z["z"]


def state_dict(prefix):
    z = golden()
    pre = prefix + "/"
    keys = [k for k in z.files if k.startswith(pre) and "/" not in k[len(pre):]]
    return {k[len(pre):]: z[k] for k in keys}


def meta():
    return json.loads(bytes(golden()["meta"]).decode())

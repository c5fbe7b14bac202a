"""Loader + oracle replay of tests/golden/brain_golden3.npz (minted by oracle/make_brain_golden3.py from the
reference's PERDQNAgent).  Used by the CPU pinning test and by the GPU parity test."""
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "brain_golden3.npz")
_z = {}


def golden3():
    if "z" not in _z:
        _z["z"] = np.load(PATH)
    return _z["z"]


def sd3(prefix):
    z = golden3()
    pre = prefix + "/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre) and "/" not in k[len(pre):]}


def meta3():
    return json.loads(bytes(golden3()["meta"]).decode())["perdqn"]


def transitions(lo, hi):
    z = golden3()
    return (z["run/state"][lo:hi], z["run/action"][lo:hi], z["run/reward"][lo:hi], z["run/next_state"][lo:hi],
            z["run/done"][lo:hi])

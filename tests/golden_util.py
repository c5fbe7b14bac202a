"""Loader for tests/golden/world_golden.npz (written by oracle/make_golden.py from the real reference)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "world_golden.npz")
_cache = {}


def load_cases():
    if "cases" not in _cache:
        z = np.load(GOLDEN)
        meta = json.loads(bytes(z["meta"]).decode())
        cases = []
        for k, m in enumerate(meta):
            c = dict(m)
            for name in ("in_type", "in_rec", "out_type", "out_rec", "out_reward", "out_obs"):
                key = f"c{k}_{name}"
                if key in z.files:
                    c[name] = z[key]
            cases.append(c)
        _cache["cases"] = cases
    return _cache["cases"]


REC_FIELDS = ("cell", "health", "age", "max_age", "gene", "flags", "action", "prev_slot")


def rec_equal(a, b, fields=REC_FIELDS):
    return all((a[f] == b[f]).all() for f in fields)


def rec_diff(a, b, fields=REC_FIELDS):
    return {f: (a[f].tolist(), b[f].tolist()) for f in fields if not (a[f] == b[f]).all()}

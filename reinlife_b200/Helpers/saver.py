"""Checkpoint writer in the reference's on-disk layout (ReinLife/Helpers/saver.py:58-97, World/entities.py:224-242):
experiments/<date>_V<n>/<METHOD>/brain_gene_<g>.pt (the eval network's state_dict, reference key names and shapes)
plus parameters_gene_<g>.json with the brain's scalar attributes, results.json (tracker.results, saver.py:84-85) and
settings.json (environment.py:243-247).  results.png needs matplotlib and is out of scope."""
import json
import os
from datetime import date

import torch


def save_brains(env, root="experiments"):
    today = str(date.today())
    v = 1                                               # <date>_V1, _V2, ... (saver.py:121-127)
    while os.path.exists(os.path.join(root, f"{today}_V{v}")):
        v += 1
    path = os.path.join(root, f"{today}_V{v}")
    out = []
    for gene, brain in enumerate(env.brains):
        d = os.path.join(path, brain.method)
        os.makedirs(d, exist_ok=True)
        net = getattr(brain, "eval_net", None) or getattr(brain, "agent", None) or getattr(brain, "model", None)
        f = os.path.join(d, f"brain_gene_{gene}.pt")
        torch.save(net.state_dict(), f)
        params = {k: v for k, v in vars(brain).items() if isinstance(v, (int, float, str, bool)) and not k.startswith("_")}
        with open(os.path.join(d, f"parameters_gene_{gene}.json"), "w") as fh:
            json.dump(params, fh, indent=4)
        out.append(f)
    results = getattr(getattr(env, "tracker", None), "results", None)
    if results is not None:
        with open(os.path.join(path, "results.json"), "w") as fh:
            json.dump(results, fh, indent=4)
    with open(os.path.join(path, "settings.json"), "w") as fh:
        json.dump({"Update interval": env.update_interval, "Width": env.width, "Height": env.height,
                   "Max agents": env.max_agents, "Families": env.static_families}, fh, indent=4)
    return out

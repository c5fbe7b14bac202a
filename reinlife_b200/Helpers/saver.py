"""Checkpoint writer with the reference's interface and on-disk layout (ReinLife/Helpers/saver.py:12-214,
World/environment.py:233-256, World/entities.py:224-242):

    <main_folder>/<date>_V<n>/<METHOD>/brain_gene_<g>.pt          static families (saver.py:130-134)
    <main_folder>/<date>_V<n>/<METHOD>/brain_<k>.pt               otherwise, k = 1.. per method (saver.py:136-145)
    .../parameters_gene_<g>.json | parameters_<k>.json            the brain's scalar attributes (saver.py:170-194)
    <main_folder>/<date>_V<n>/results.json, settings.json         tracker.results, environment settings (saver.py:84-88)

`.pt` files hold the state_dict of the network the reference saves for each method (eval_net / agent / model) under the
reference's key names and shapes, so they load with the reference's `load_model=` and vice versa.  Differences, on
purpose: paths are joined with os.path.join (the reference hard-codes "\\\\" unless google_colab, saver.py:55, which on
Linux creates directories with backslashes in their names); `fig` is accepted and saved only if it has `savefig`
(matplotlib is out of scope here)."""
import inspect
import json
import os
from datetime import date

import torch


class _SavedAgent:
    """What Saver.save iterates over: the reference passes Agent objects (environment.py:252-256); only `.brain` and
    `.gene` are read."""

    def __init__(self, gene, brain):
        self.gene, self.brain = gene, brain


def _network_of(brain):
    """Agent.save_brain's dispatch on brain.method (entities.py:224-242)."""
    method = brain.method
    if method == "DQN":
        return brain.agent
    if method in ("PERDQN", "PPO"):
        return brain.model
    if method in ("PERD3QN", "DRQN", "D3QN"):
        return brain.eval_net
    raise ValueError(f"no checkpoint rule for brain method {method!r}")


class Saver:
    """Saver(main_folder, google_colab=False).save(agents, family, results, settings, fig) -- saver.py:51-97."""

    def __init__(self, main_folder: str, google_colab: bool = False):
        self.google_colab = google_colab
        self.separator = os.sep
        self.main_folder = os.path.join(os.getcwd(), main_folder)

    def save(self, agents, family: bool, results: dict, settings: dict, fig=None):
        exp = self._experiment_dir()
        stems = self._brain_stems(agents, family, exp)
        for d in sorted({os.path.dirname(stem) for stem in stems.values()}):
            os.makedirs(d, exist_ok=True)
        written = []
        for agent, stem in stems.items():
            torch.save(_network_of(agent.brain).state_dict(), stem + ".pt")
            written.append(stem + ".pt")
            folder, base = os.path.split(stem)
            with open(os.path.join(folder, base.replace("brain", "parameters") + ".json"), "w") as f:
                json.dump(self._scalar_attributes(agent.brain), f, indent=4)
        for name, payload in (("results.json", results), ("settings.json", settings)):
            with open(os.path.join(exp, name), "w") as f:
                json.dump(payload, f, indent=4)
        if fig is not None and hasattr(fig, "savefig"):
            fig.savefig(os.path.join(exp, "results.png"), dpi=150)
        return written

    def _experiment_dir(self):
        """<main_folder>/<today>_V<n>, n = 1 + the highest version already present for today (saver.py:119-127)."""
        os.makedirs(self.main_folder, exist_ok=True)
        tag = f"{date.today()}_V"
        taken = [int(name[len(tag):]) for name in os.listdir(self.main_folder)
                 if name.startswith(tag) and name[len(tag):].isdigit()]
        exp = os.path.join(self.main_folder, tag + str(max(taken, default=0) + 1))
        os.makedirs(exp)
        return exp

    @staticmethod
    def _brain_stems(agents, family, exp):
        """agent -> path without extension: brain_gene_<gene> for static families (saver.py:130-134), otherwise
        brain_1, brain_2, ... per method in list order (saver.py:136-145)."""
        stems, per_method = {}, {}
        for agent in agents:
            method = agent.brain.method
            if family:
                name = f"brain_gene_{agent.gene}"
            else:
                per_method[method] = per_method.get(method, 0) + 1
                name = f"brain_{per_method[method]}"
            stems[agent] = os.path.join(exp, method, name)
        return stems

    @staticmethod
    def _scalar_attributes(brain):
        """saver.py:170-194: every non-routine member whose type is exactly float / int / bool / str.  Class constants
        of the device brains (KIND, RULE, ...) and private state other than `_method` are left out."""
        out = {}
        for name, val in inspect.getmembers(brain, lambda a: not inspect.isroutine(a)):
            if type(val) in (float, int, bool, str) and not name.isupper() and (not name.startswith("_") or name == "_method"):
                out[name] = val
        return out


def save_brains(env, root="experiments"):
    """Environment.save_results (environment.py:233-256): one stand-in agent per brain for static families."""
    if hasattr(env, "sync_host_scalars"):
        env.sync_host_scalars()            # epsilon / n_epi live on the device while training
    settings = {"Update interval": env.update_interval, "Width": env.width, "Height": env.height,
                "Max agents": env.max_agents, "Families": env.static_families}
    results = getattr(getattr(env, "tracker", None), "results", None)
    agents = [_SavedAgent(gene, brain) for gene, brain in enumerate(env.brains)]
    return Saver(root, google_colab=env.google_colab).save(agents, env.static_families, results, settings, None)

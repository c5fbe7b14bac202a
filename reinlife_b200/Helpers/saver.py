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
    def __init__(self, main_folder: str, google_colab: bool = False):
        self.google_colab = google_colab
        self.separator = os.sep
        self.main_folder = os.path.join(os.getcwd(), main_folder)

    # -- saver.py:58-97
    def save(self, agents, family: bool, results: dict, settings: dict, fig=None):
        directory_paths, agent_paths, experiment_path = self._get_paths(agents, family)
        self._create_directories(directory_paths)
        written = []
        for agent in agents:
            torch.save(_network_of(agent.brain).state_dict(), agent_paths[agent] + ".pt")
            written.append(agent_paths[agent] + ".pt")
        with open(os.path.join(experiment_path, "results.json"), "w") as f:
            json.dump(results, f, indent=4)
        with open(os.path.join(experiment_path, "settings.json"), "w") as f:
            json.dump(settings, f, indent=4)
        if fig is not None and hasattr(fig, "savefig"):
            fig.savefig(os.path.join(experiment_path, "results.png"), dpi=150)
        self._save_params(agents, agent_paths)
        return written

    # -- saver.py:99-147
    def _get_paths(self, agents, family: bool):
        today = str(date.today())
        experiment_path = os.path.join(self.main_folder, today + "_V1")
        if os.path.exists(experiment_path):
            paths = [path for path in os.listdir(self.main_folder) if today in path]
            index = str(max(self.get_int(path.split("V")[-1]) for path in paths) + 1)
            experiment_path = experiment_path[:-1] + index
        model_paths = sorted({os.path.join(experiment_path, agent.brain.method) for agent in agents})
        if family:
            agents_paths = {agent: os.path.join(experiment_path, agent.brain.method, "brain_gene_" + str(agent.gene))
                            for agent in agents}
        else:                       # brain_1, brain_2, ... per method, in the order the agents are listed
            agents_paths, seen = {}, {}
            for agent in agents:
                seen[agent.brain.method] = seen.get(agent.brain.method, 0) + 1
                agents_paths[agent] = os.path.join(experiment_path, agent.brain.method, "brain_" + str(seen[agent.brain.method]))
        return [self.main_folder, experiment_path] + model_paths, agents_paths, experiment_path

    def _create_directories(self, all_paths):
        for path in all_paths:
            if not os.path.exists(path) and not self._create_directory(path):
                raise Exception(f"{path} could not be created")

    # -- saver.py:170-194: every non-routine member whose type is exactly float / int / bool / str
    @staticmethod
    def _save_params(agents, agent_paths):
        for agent in agents:
            params = {}
            for name, val in inspect.getmembers(agent.brain, lambda a: not inspect.isroutine(a)):
                if type(val) in (float, int, bool, str) and not name.isupper() and (not name.startswith("_") or name == "_method"):
                    params[name] = val          # class constants of the device brains (KIND, RULE, ...) and private state are skipped
            folder, base = os.path.split(agent_paths[agent])
            with open(os.path.join(folder, base.replace("brain", "parameters") + ".json"), "w") as f:
                json.dump(params, f, indent=4)

    @staticmethod
    def _create_directory(path: str) -> bool:
        try:
            os.mkdir(path)
        except OSError:
            return False
        return True

    @staticmethod
    def get_key(val, dictionary):
        return next(key for key, value in dictionary.items() if value == val)

    @staticmethod
    def get_int(a_string: str) -> int:
        return int("".join(s for s in a_string if s.isdigit()))


def save_brains(env, root="experiments"):
    """Environment.save_results (environment.py:233-256): one stand-in agent per brain for static families."""
    settings = {"Update interval": env.update_interval, "Width": env.width, "Height": env.height,
                "Max agents": env.max_agents, "Families": env.static_families}
    results = getattr(getattr(env, "tracker", None), "results", None)
    agents = [_SavedAgent(gene, brain) for gene, brain in enumerate(env.brains)]
    return Saver(root, google_colab=env.google_colab).save(agents, env.static_families, results, settings, None)

"""Tracker -- the reference's per-gene statistics (ReinLife/Helpers/tracker.py:57-78,107-132,178-282) fed from device
reductions.  Per step the stats kernel writes one record into a device ring; the ring crosses to the host only
every `update_interval` steps (the reference averages at the same cadence, tracker.py:121), so the step loop stays
free of host syncs.  With N worlds the per-step value of each series is pooled over all worlds; at N=1 it is the
reference's value, including its quirks ("Avg Number of Kills" appends the SUM of agent.killed, tracker.py:256;
"Avg Number of Intra Kills" is 1 whenever there was a kill, :252-261; -1 marks "no agents", filtered by :280)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib

VARIABLES = ["Avg Population Size", "Avg Population Age", "Avg Population Fitness", "Best Population Age",
             "Avg Number of Attacks", "Avg Number of Kills", "Avg Number of Intra Kills", "Avg Number of Populations"]


def series_from_record(rec, n_genes):
    """One stats record of rl_world_stats -> the per-step value of every series, as Tracker._track_results appends them
    (tracker.py:178-266).  Index order per gene: size, age, fitness, best age, attacks, kills, intra kills."""
    G, NS = n_genes, _lib.N_STATS
    out = {}
    tail = rec[G * NS:]
    if tail[0] == 0:              # no agent anywhere: the reference appends -1 to EVERY series (tracker.py:189-199)
        for g in range(G):
            out[g] = [-1] * 7
        out["populations"] = -1
        return out
    for g in range(G):
        cnt, age, rew, amax, att, kil, worlds = rec[g * NS:g * NS + 7]
        if cnt == 0:              # other genes alive: sum([]) = 0 kills, 0 intra kills (tracker.py:252-261)
            vals = [-1, -1, -1, -1, -1, 0.0, 0]
        else:
            vals = [cnt / worlds, age / cnt, rew / cnt, amax, att / cnt, kil / max(worlds, 1.0), 1.0 if kil != 0 else 0]
        out[g] = vals
    out["populations"] = tail[2] / tail[1] if tail[1] > 0 else -1
    return out


def average_rows(rows, n_genes):
    """Tracker._aggregate (tracker.py:279-282) over the rows of one update interval: mean of the values > -1."""
    res = {}
    for vi, var in enumerate(VARIABLES[:-1]):
        res[var] = []
        for g in range(n_genes):
            vals = [r[g][vi] for r in rows if r[g][vi] > -1]
            res[var].append(float(np.mean(vals)) if vals else float("nan"))
    vals = [r["populations"] for r in rows if r["populations"] > -1]
    res[VARIABLES[-1]] = float(np.mean(vals)) if vals else float("nan")
    return res


class Tracker:
    def __init__(self, env, update_interval, print_results=True):
        self.env, self.update_interval, self.print_results = env, int(update_interval), print_results
        self.nr_genes = len(env.brains)
        G = self.nr_genes
        self.results = {v: ({g: [] for g in range(G)} if v != VARIABLES[-1] else []) for v in VARIABLES}
        self.variables = list(self.results.keys())
        w = env.world
        self.nv = G * _lib.N_STATS + 8
        self.ring_len = max(1, min(self.update_interval, 4096))
        dev = env.device
        self.ring = torch.zeros((self.ring_len, self.nv), dtype=torch.float64, device=dev)
        n_scr = w.lib.rl_world_stats_scratch_doubles(C.byref(w.cfg))
        self.scratch = torch.zeros(n_scr, dtype=torch.float64, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        # step stamp echoed by the stats kernel: one pinned control block PER ring slot, so that a host that runs ahead of
        # the device never rewrites a block whose H2D copy is still pending (the drain below syncs every ring_len steps)
        self.ctrl = torch.zeros((self.ring_len, 2), dtype=torch.int64, device=dev)
        self.ctrl_host = torch.zeros((self.ring_len, 2), dtype=torch.int64).pin_memory()
        self.k = 0
        self.rows_host = []          # per-step series values since the last aggregation
        self.history = None          # set to [] to keep every per-step series (parity tests: tracker.track_results)
        self.fig = None

    def record(self, n_epi):
        """Launch the reduction for the current agent list into the next ring slot (no sync)."""
        w = self.env.world
        slot = self.k % self.ring_len
        self.ctrl_host[slot, 0] = n_epi
        self.ctrl[slot].copy_(self.ctrl_host[slot], non_blocking=True)
        out = self.ring[slot]
        with torch.cuda.device(self.env.device):
            _lib.check(w.lib.rl_world_stats(C.byref(w.cfg), C.byref(w.bufs), C.c_void_p(self.ctrl[slot].data_ptr()),
                                            C.c_void_p(self.scratch.data_ptr()), C.c_void_p(self.counter.data_ptr()),
                                            C.c_void_p(out.data_ptr()), w._stream()))
        self.k += 1
        if self.k % self.ring_len == 0:
            self._drain(self.ring_len)
        return out

    def _drain(self, n):
        if n <= 0:
            return
        host = self.ring[:n].cpu().numpy()
        for rec in host:
            row = self.series_from_record(rec)
            self.rows_host.append(row)
            if self.history is not None:
                self.history.append(row)

    def series_from_record(self, rec):
        return series_from_record(rec, self.nr_genes)

    def update_results(self, agents=None, n_epi=0):
        """Same cadence as the reference (tracker.py:107-132): called every step from update_env."""
        self.record(n_epi)
        if n_epi % self.update_interval == 0 and n_epi != 0:
            self._drain(self.k % self.ring_len)
            self.k = 0
            self.env.check_status()
            self.env.sync_host_scalars()
            self._average_results()
            if self.print_results:
                self._print_results()

    def _average_results(self):
        res = average_rows(self.rows_host[-self.update_interval:], self.nr_genes)
        for var in VARIABLES[:-1]:
            for g in range(self.nr_genes):
                self.results[var][g].append(res[var][g])
        self.results[VARIABLES[-1]].append(res[VARIABLES[-1]])
        self.rows_host = []

    def _print_results(self):
        print(f"################ {self.env.n_worlds_global} worlds ################")
        for var in VARIABLES[:-1]:
            print(f"{var:28s}" + "".join(f"  gene {g}: {self.results[var][g][-1]:8.3f}" for g in range(self.nr_genes)))
        print(f"{VARIABLES[-1]:28s}  {self.results[VARIABLES[-1]][-1]:8.3f}")

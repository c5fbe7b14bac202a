"""Turns the raw ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/make_profile_summary.py gpurun_out/launches_r01d.csv gpurun_out/full_r01d.ncu-rep r01
"""
import collections
import csv
import json
import os
import subprocess
import sys

launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
# the ncu selection of the full capture (4th argument), as written into the summary header
FULL_SEL = sys.argv[4] if len(sys.argv) > 4 else ('-k regex:"k_learn_dueling_p|k_world_step|k_world_update|k_replay_store|k_act_dueling_p|k_replay_sample|'
                                                  'k_replay_update_prio|k_world_stats" -s 40 -c 12')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")

rows = list(csv.reader(open(launch_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi or not r[vi]:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v = v / 1e3 if r[ui] == "us" else v / 1e6 if r[ui] == "ns" else v
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out_dir, f"launches_{tag}_summary.md"), "w") as f:
    f.write(f"# ncu launch list, {tag}\n\nCommand (under gpurun, 1x B200): `ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv "
            f"python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\nRaw list: `profiles/{os.path.basename(launch_csv)}`. "
            "Times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
            f"Total device time over {sum(a[0] for a in agg.values())} launches: {tot:.1f} ms\n\n"
            "| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k[:80]}` | {n} | {t:.3f} | {t / tot * 100:.1f}% | {t / n:.4f} |\n")

metrics = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(metrics)], text=True)
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
seen, lines, traffic = set(), [], {}
cols = [c for c in hdr if c in metrics]
for row in r[2:]:
    d = dict(zip(hdr, row))
    name = d["Kernel Name"].replace("<unnamed>::", "").replace("void ", "").split("(")[0].split("<")[0]
    if name in seen:
        continue
    seen.add(name)
    lines.append("| " + name + " | " + " | ".join(d[c] for c in cols) + " |")
    u = dict(zip(hdr, units))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    traffic[name + "_bytes_per_launch"] = int(float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]] +
                                              float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]])
with open(os.path.join(out_dir, f"ncu_full_{tag}_summary.md"), "w") as f:
    f.write(f"# ncu --set full, {tag} (one row per kernel; first captured launch)\n\n"
            f"Command: `ncu --set full --clock-control none --import-source on {FULL_SEL} "
            "python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (the event kernel's name is k_learn_dueling_h in the round-1 captures)\n\n| kernel | " + " | ".join(cols) + " |\n|" + "---|" * (len(cols) + 1) + "\n"
            "| (unit) | " + " | ".join(dict(zip(hdr, units))[c] for c in cols) + " |\n" + "\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(out_dir, f"traffic_{tag}.json"), "w"), indent=1)
print(open(os.path.join(out_dir, f"ncu_full_{tag}_summary.md")).read())

"""Turns the ncu CSV exports of a gpurun call into the tracked evidence files under profiles/ (run in the build container):
    python scripts/summarize_profiles.py gpurun_out/c15_launches_raw.csv gpurun_out/c15_full_raw.csv r02
-> profiles/{tag}_launches_summary.csv (per-kernel share of the timed steps), profiles/{tag}_ncu_full_summary.csv (one line per
captured kernel: duration, DRAM bytes, pipe utilisation), profiles/{tag}_traffic.json (DRAM bytes per launch, read by bench.py)."""
import collections
import csv
import json
import shutil
import sys

launch_csv, full_csv, tag = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(launch_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) > mv:
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("roreg::", "")
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(t for _, t in agg.values())
with open(f"profiles/{tag}_launches_summary.csv", "w") as f:
    f.write(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none) of the default bench (pipelined schedule, score mode 1, nn mode 4, corr mode 3, 64 pairs per step)\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's stage_ms_per_step\n")
    f.write("kernel,launches,total_us,share_pct,avg_us\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{t / 1e3:.1f},{100 * t / tot:.1f},{t / n / 1e3:.1f}\n")
shutil.copy(launch_csv, f"profiles/{tag}_launches_raw.csv")

rows = list(csv.reader(open(full_csv)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic"]
cols = []
for w in want:
    for i, x in enumerate(hdr):
        if x == w or x.endswith("." + w) or x.endswith(w) and w != "Kernel Name":
            cols.append((w, i)); break
    else:
        if w == "Kernel Name":
            cols.append((w, hdr.index(w)))
units = rows[1]
traffic = {}
with open(f"profiles/{tag}_ncu_full_summary.csv", "w") as f:
    f.write(f"# {tag}: ncu --set full --clock-control none, one launch per kernel of a 64-pair bench step (units: " + "; ".join(f"{w}={units[i]}" for w, i in cols if units[i]) + ")\n")
    f.write(",".join(w for w, _ in cols) + "\n")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        f.write(",".join('"' + r[i].split("(")[0].replace("void ", "") + '"' if w == "Kernel Name" else r[i] for w, i in cols) + "\n")
        d = dict((w, r[i]) for w, i in cols); u = dict((w, units[i]) for w, i in cols)

        def to_bytes(key):
            v = float(d[key]); s = u[key].lower()
            return v * (1e9 if s.startswith("g") else 1e6 if s.startswith("m") else 1e3 if s.startswith("k") else 1.0)
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("roreg::", "").split("<")[0]
        traffic[name] = {"dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"), "pairs_per_launch": 64}
json.dump({"source": f"profiles/{tag}_ncu_full_summary.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, 64 pairs per launch)",
           "kernels": traffic}, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print(open(f"profiles/{tag}_launches_summary.csv").read())
print(open(f"profiles/{tag}_ncu_full_summary.csv").read()[:3000])

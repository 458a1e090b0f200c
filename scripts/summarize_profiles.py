"""Turns the ncu artefacts of scripts/make_profiles.sh (gpurun_out/) into the committed summaries
under profiles/: per-kernel launch list, key metrics of the full capture, traffic.json."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
shape = os.environ.get("PROBE_SHAPE", "512,512,512")
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")

# ---- launch list
lines = [l for l in open(os.path.join(go, f"launches_{tag}.csv")) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0][:60]
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else v * 1e3 if row["Metric Unit"] == "ms" else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out, f"{tag}_launches.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: 3 RL iterations at {shape} (slices,H,W) incl. OTF preparation\n")
    f.write("# cold-cache, serialised per-launch times: compare SHARES, not absolutes (scripts/make_profiles.sh)\n")
    f.write(f"{'kernel':60s} {'n':>5s} {'total us':>10s} {'avg us':>9s} {'share':>7s}\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:60s} {n:5d} {t:10.1f} {t / n:9.1f} {100 * t / tot:6.1f}%\n")

# ---- full capture
def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    return rows[0], rows[1], rows[2:]

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle"]
traffic = {}
for rep, label in ((f"prof_{tag}.ncu-rep", "rl_iteration"), (f"prof_zncc_{tag}.ncu-rep", "zncc")):
    path = os.path.join(go, rep)
    if not os.path.exists(path):
        continue
    hdr, units, data = raw(path)
    ki = hdr.index("Kernel Name")
    with open(os.path.join(out, f"{tag}_ncu_{label}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({rep}); one column per profiled launch\n")
        f.write(f"{'metric':62s} {'unit':>10s} " + " ".join(f"{r[ki].split('(')[0][5:19]:>14s}" for r in data) + "\n")
        for w in WANT:
            if w not in hdr:
                continue
            i = hdr.index(w)
            f.write(f"{w[:62]:62s} {units[i][:10]:>10s} " + " ".join(f"{r[i][:14]:>14s}" for r in data) + "\n")
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    def tobytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    per = [(r[ki].split("(")[0][5:], tobytes(r[ir], units[ir]) + tobytes(r[iw], units[iw]), float(r[it])) for r in data]
    if label == "rl_iteration":
        traffic["dram_bytes_per_iteration"] = sum(p[1] for p in per[:8])
        traffic["kernels"] = [{"kernel": p[0][:40], "dram_bytes": p[1], "us_under_ncu": p[2]} for p in per[:8]]
    else:
        traffic["zncc_dram_bytes_per_evaluation"] = per[0][1]
traffic["source"] = f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, {tag}; {shape} (slices,H,W) single view; zncc at 512x512x256"
json.dump(traffic, open(os.path.join(out, "traffic.json"), "w"), indent=1)
print(open(os.path.join(out, f"{tag}_launches.txt")).read())
print(json.dumps(traffic, indent=1)[:1500])

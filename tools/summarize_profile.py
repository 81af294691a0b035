"""Turns ncu outputs (gpurun_out/*.ncu-rep via `ncu -i --page raw --csv`, and a launch-list csv) into the
committed summaries under profiles/.  usage: summarize_profile.py <tag> <launches.csv> <full.ncu-rep> [workload]"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
workload = sys.argv[4] if len(sys.argv) > 4 else "c1"
out = [f"# ncu summary {tag}\n"]
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]; i_name = hdr.index("Kernel Name"); i_val = hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[i_val].replace(",", ""))
    except ValueError: continue
    k = r[i_name].split("(")[0]; agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
out.append(f"## Launch list ({os.path.basename(launches)}: `ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised)\n")
out.append("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| {k} | {n} | {t/1000:.1f} | {100*t/tot:.1f} % |")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]; idx = {x: i for i, x in enumerate(h)}
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]
out.append(f"\n## Full capture ({os.path.basename(rep)}: `ncu --set full --clock-control none --import-source on`)\n")
traffic = {}
for r in rr[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    out.append(f"### {name}\n")
    out.append("| metric | value | unit |\n|---|---:|---|")
    for w in want:
        if w in idx:
            out.append(f"| {w} | {r[idx[w]]} | {rr[1][idx[w]]} |")
    try:
        def b(k):
            v = float(r[idx[k]]); u = rr[1][idx[k]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        key = name.replace("void ", "").split("<")[0]
        traffic[key] = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
        out.append(f"| dram traffic per launch | {traffic[key]/1e6:.2f} | MB |")
    except (KeyError, ValueError): pass
    out.append("")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", tag + ".md"), "w").write("\n".join(out) + "\n")
tp = os.path.join(ROOT, "profiles", "traffic.json")
allt = json.load(open(tp)) if os.path.exists(tp) else {}
allt[workload] = {k: round(v) for k, v in traffic.items()}
json.dump(allt, open(tp, "w"), indent=1)
print("\n".join(out[:40]))

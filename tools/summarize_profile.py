"""Turns ncu outputs (a `--set full` .ncu-rep and, optionally, a launch-list csv) into the committed summaries under
profiles/: <tag>.md (tables) and ncu_final.json (the per-kernel counters bench.py attaches to its roofline record:
issue_frac, l2_gbs, lane_efficiency, dram_bytes of ONE steady-state launch).
usage: summarize_profile.py <tag> <full.ncu-rep> <workload> [launches.csv]"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep, workload = sys.argv[1], sys.argv[2], sys.argv[3]
launches = sys.argv[4] if len(sys.argv) > 4 else None
out = [f"# ncu summary {tag} ({workload})\n"]
if launches:
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
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
unit_scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3,
              "nsecond": 1e-9, "second": 1, "s": 1}
def val(r, k):
    v = float(r[idx[k]].replace(",", "")); u = rr[1][idx[k]].lower()
    return v * unit_scale.get(u, 1)
def l2_bytes(r):
    """lts__t_bytes when the capture has it (--metrics lts__t_bytes.sum next to --set full), else the L1 <-> L2 crossbar bytes."""
    if "lts__t_bytes.sum" in idx:
        return val(r, "lts__t_bytes.sum")
    return val(r, "l1tex__m_xbar2l1tex_read_bytes.sum") + (val(r, "l1tex__m_l1tex2xbar_write_bytes.sum") if "l1tex__m_l1tex2xbar_write_bytes.sum" in idx else 0.0)
out.append(f"\n## Full capture ({os.path.basename(rep)}: `ncu --set full --clock-control none --import-source on`)\n")
facts = {}
for r in rr[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    out.append(f"### {name}\n")
    out.append("| metric | value | unit |\n|---|---:|---|")
    for w in want:
        if w in idx:
            out.append(f"| {w} | {r[idx[w]]} | {rr[1][idx[w]]} |")
    try:
        key = name.replace("void ", "").replace("crt::", "").split("<")[0]
        dur = val(r, "gpu__time_duration.sum")
        if key in facts and facts[key]["duration_ms"] >= round(dur * 1e3, 4):
            out.append("")
            continue                                      # several captures of one kernel: the longest launch is the one kept
        facts[key] = {
            "issue_frac": round(float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]) / 100, 4),
            "lane_efficiency": round(float(r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]]) / 32, 4),
            "l2_gbs": round(l2_bytes(r) / dur / 1e9, 1),
            "dram_bytes": int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")),
            "duration_ms": round(dur * 1e3, 4),
            "pipe_alu_frac": round(float(r[idx["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]]) / 100, 4),
            "pipe_fma_frac": round(float(r[idx["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]]) / 100, 4),
            "warp_inst": int(float(r[idx["smsp__inst_executed.sum"]])),
            "source": f"profiles/{tag}.md ({os.path.basename(rep)})",
        }
        out.append(f"| dram traffic per launch | {facts[key]['dram_bytes']/1e6:.2f} | MB |")
        out.append(f"| L2 traffic rate (lts__t_bytes / duration) | {facts[key]['l2_gbs']:.1f} | GB/s |")
    except (KeyError, ValueError) as e:
        out.append(f"| (facts not derived: {e}) | | |")
    out.append("")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", tag + ".md"), "w").write("\n".join(out) + "\n")
tp = os.path.join(ROOT, "profiles", "ncu_final.json")
allt = json.load(open(tp)) if os.path.exists(tp) else {}
allt[workload] = facts
json.dump(allt, open(tp, "w"), indent=1)
print("\n".join(out[-60:]))

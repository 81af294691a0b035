"""Per-CUDA-source-line totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:NAME` (development aid).
usage: ncu_lines.py dump.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = "?"; cur = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if r[0] == "Kernel Name": print(r[1][:100]); continue
    if hdr is None: continue
    if r[0] != "":
        cur = (fname, int(r[0]) if r[0].isdigit() else -1, r[1].strip()[:90])
        agg.setdefault(cur, [0.0, 0.0, 0.0, collections.Counter()])
        continue
    if cur is None or len(r) < len(hdr) - 2 or r[2] in ("...", ""): continue
    def f(k):
        try: return float(r[idx[k]])
        except (ValueError, KeyError, IndexError): return 0.0
    a = agg[cur]
    a[0] += f("# Samples"); a[1] += f("Instructions Executed"); a[2] += f("Thread Instructions Executed")
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            v = f(k)
            if v: a[3][k[6:]] += v
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print("total samples %d, warp instructions %d, lanes/inst %.1f" % (ts, ti, sum(a[2] for a in agg.values()) / max(ti, 1)))
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    st = " ".join("%s:%d" % (n, v) for n, v in a[3].most_common(3))
    print("%5.1f%%s %5.1f%%i  %s:%d  %-90s %s" % (100 * a[0] / ts, 100 * a[1] / ti, k[0], k[1], k[2], st))

"""Top stall sites per SASS instruction from `ncu -i X.ncu-rep --page source --csv --print-source sass -k regex:NAME` (development aid).
usage: ncu_sass_top.py dump.csv [top_n] [context]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hdr = next(r for r in rows if r and r[0] == "Address")
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
S = idx["# Samples"]
def fl(x):
    try: return float(x)
    except ValueError: return 0.0
tot = sum(fl(r[S]) for r in data)
print("total samples", tot, "instructions", len(data))
top = sorted(range(len(data)), key=lambda i: -fl(data[i][S]))[:top_n]
shown = set()
for i in sorted(top):
    for j in range(max(0, i - ctx), i + 1):
        if j in shown: continue
        shown.add(j)
        r = data[j]
        st = {k[6:]: int(fl(r[idx[k]])) for k in hdr if k.startswith("stall_") and "Not" not in k and fl(r[idx[k]]) > 0}
        st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
        print("%s%4d %5.1f%% exec %10s thr %5s | %-72s | %s" % ("*" if j == i else " ", j, 100 * fl(r[S]) / tot, r[idx["Instructions Executed"]],
                                                          r[idx["Avg. Threads Executed"]], r[1].strip()[:72], st))

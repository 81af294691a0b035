#!/bin/bash
# round 2, session 35: k_shade alone (ncu, one kernel at a time) with and without the zero-numerator guard on the same box - the
# launch lists of s32 / s34 show it at 2.0-2.2 ms per launch where the list of s23 (binary before the guard) had 1.46 ms, while the
# un-profiled frames got faster
mkdir -p gpurun_out
for v in variants/libcrt_nodiv.so libcrt.so; do
  echo "== $v"
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=3840 QB_H=2160 QB_SPP=24 QB_NO_BATCH=1 QB_SCENES=cornell-box timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_shade|k_generate' -c 400 --csv --log-file gpurun_out/s35_$(basename $v .so).csv python tools/quick_bench.py ploc8 | grep "est 0"
  python - <<PY
import csv, statistics
rows = list(csv.reader(open("gpurun_out/s35_$(basename $v .so).csv")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
idx = {x: i for i, x in enumerate(rows[h])}
d = {}
for r in rows[h + 1:]:
    if len(r) < len(rows[h]): continue
    v = float(r[idx["Metric Value"]].replace(",", ""))
    u = r[idx["Metric Unit"]]
    v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
    d.setdefault(r[idx["Kernel Name"]].split("(")[0], []).append(v)
for k, v in d.items():
    print(k, len(v), "median %.0f us" % statistics.median(v), [round(x) for x in v[20:32]])
PY
done 2>&1 | tee gpurun_out/r02_s35.log

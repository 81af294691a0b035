#!/bin/bash
# round 2, final evidence: ncu launch list of the default bench command and full captures (with source) of the final binary -
# k_extend / k_shade / k_shadow at C3 steady state (2^25 paths in flight), the C5 batch kernel, the tail path tracer on C1.
# Summaries: python tools/summarize_profile.py <tag> <rep> <workload> [launches.csv]  ->  profiles/<tag>.md + profiles/ncu_final.json
mkdir -p gpurun_out
EXTRA=lts__t_bytes.sum,lts__t_sectors.sum
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 1 --no-sub > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-120
timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade' -s 60 -c 3 -o gpurun_out/r02_final_c3 -f python bench.py --steps 1 --warmup 0 --no-sub > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-120
CRT_C5_RAYS=20000000 CRT_C5_E2E_RAYS=1000000 timeout 900 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'k_trace_batch' -s 1 -c 1 -o gpurun_out/r02_final_c5 -f python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/ncu_c5.log 2>&1
tail -1 gpurun_out/ncu_c5.log | cut -c1-120
timeout 600 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:'k_tail|k_extend|k_shadow' -s 12 -c 3 -o gpurun_out/r02_final_c1 -f python tools/c1_timeline.py cornell-box > gpurun_out/ncu_c1.log 2>&1
tail -1 gpurun_out/ncu_c1.log | cut -c1-120
ls -la gpurun_out/*.ncu-rep

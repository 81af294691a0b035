#!/bin/bash
# round 2: compute-sanitizer over every kernel family (SURVEY.md section 5 "race detection": the wavefront replaces the
# reference's race-free one-thread-per-pixel kernel, Render.cuh:332-336, with global atomics, __syncwarp-ordered shared queues
# and one intended race, the any-hit 64-bit store). Logs -> gpurun_out/r02_sanitizer_*.log, summary in profiles/r02_sanitizer.md.
mkdir -p gpurun_out
export CRT_POOL=8192
for tool in memcheck initcheck; do
  SAN_SCENES=veach-mis,cornell-box timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -c '=========' gpurun_out/r02_sanitizer_$tool.log) lines"; grep -E "ERROR SUMMARY|done$" gpurun_out/r02_sanitizer_$tool.log | tail -3
done
SAN_SCENES=veach-mis SAN_W=32 SAN_H=24 timeout 2400 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 40 python tools/sanitize_case.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "== racecheck"; grep -E "RACECHECK SUMMARY|done$" gpurun_out/r02_sanitizer_racecheck.log | tail -3
grep -E "Race reported|hazard" gpurun_out/r02_sanitizer_racecheck.log | sed -E 's/0x[0-9a-f]+//g' | sort | uniq -c | sort -rn | head -20

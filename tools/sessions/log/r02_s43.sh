#!/bin/bash
# round 2, session 43: compute-sanitizer initcheck and racecheck of the final binary
mkdir -p gpurun_out
export CRT_POOL=8192
SAN_SCENES=veach-mis,cornell-box timeout 120 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanitize_case.py > gpurun_out/r02_sanitizer_initcheck_final.log 2>&1
echo "== initcheck"; grep -E "ERROR SUMMARY|done$" gpurun_out/r02_sanitizer_initcheck_final.log | tail -3
SAN_SCENES=veach-mis SAN_W=32 SAN_H=24 timeout 170 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 40 python tools/sanitize_case.py > gpurun_out/r02_sanitizer_racecheck_final.log 2>&1
echo "== racecheck"; grep -E "RACECHECK SUMMARY|done$" gpurun_out/r02_sanitizer_racecheck_final.log | tail -3
grep -E "Race reported|hazard" gpurun_out/r02_sanitizer_racecheck_final.log | sed -E 's/0x[0-9a-f]+//g' | sort | uniq -c | sort -rn | head -8

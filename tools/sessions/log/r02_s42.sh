#!/bin/bash
# round 2, session 42: compute-sanitizer memcheck of the final binary (k_shade division guard, gated any-hit batches, banded ray batches)
mkdir -p gpurun_out
export CRT_POOL=8192
SAN_SCENES=veach-mis,cornell-box timeout 280 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py > gpurun_out/r02_sanitizer_memcheck_final.log 2>&1
echo "== memcheck: $(grep -c '=========' gpurun_out/r02_sanitizer_memcheck_final.log) lines"; grep -E "ERROR SUMMARY|done$" gpurun_out/r02_sanitizer_memcheck_final.log | tail -3

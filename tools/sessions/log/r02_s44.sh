#!/bin/bash
# round 2, session 44: CRT_RAY_SORTED - the ordered ray batches as an option of the product: test (same hits), cost on C5 (20 M rays)
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_synthetic.py -m gpu -x -q 2>&1 | tail -12 ) | tee gpurun_out/r02_s44.log
SP_ONLY_UNORDERED=1 timeout 200 python tools/sort_probe.py 2>&1 | tee -a gpurun_out/r02_s44.log

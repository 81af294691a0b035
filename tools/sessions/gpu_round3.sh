#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/tail_sweep.py 2>&1 | tee gpurun_out/tail_sweep.log
timeout 600 python tools/pool_sweep.py 2>&1 | tee gpurun_out/pool_sweep.log

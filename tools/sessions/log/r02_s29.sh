#!/bin/bash
# round 2, session 29 (2 GPUs): frame times of crt_group frame by frame; group + viewer tests
mkdir -p gpurun_out
( timeout 600 python tools/group_probe.py
  GP_W=800 GP_H=600 GP_SPP=2 timeout 600 python tools/group_probe.py ) 2>&1 | tee gpurun_out/r02_s29.log
( timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_viewer.py -m gpu -x -q 2>&1 | tail -30 ) | tee -a gpurun_out/r02_s29.log

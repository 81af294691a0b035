#!/bin/bash
# round 2, session 26: what would ordered rays buy on C5 (ordering done by torch outside the timed call)
mkdir -p gpurun_out
timeout 600 python tools/sort_probe.py 2>&1 | tee gpurun_out/r02_s26.log

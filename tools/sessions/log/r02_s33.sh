#!/bin/bash
# round 2, session 33: chunk size of the host-buffer ray batches (20 M and 100 M rays), the plain H2D rate of the box
mkdir -p gpurun_out
( timeout 600 python tools/e2e_probe.py; EP_RAYS=100000000 EP_CHUNKS=2097152,4194304 timeout 600 python tools/e2e_probe.py ) 2>&1 | tee gpurun_out/r02_s33.log

#!/bin/bash
# round 2: extended fuzz of the long-buffer paths
mkdir -p gpurun_out
timeout 420 python tools/gpu_fuzz_buffers.py 240 0 > gpurun_out/r02ag_fuzz.log 2>&1
tail -14 gpurun_out/r02ag_fuzz.log

#!/bin/bash
# round 2: extended fuzz on the GPU (more seeds, K5 forced on every other seed)
mkdir -p gpurun_out
timeout 400 python tools/gpu_fuzz_more.py 200 > gpurun_out/r02ae_fuzz.log 2>&1
tail -15 gpurun_out/r02ae_fuzz.log

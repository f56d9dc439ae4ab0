#!/bin/bash
# round 2: fuzz of the split-buffer search on 2 GPUs
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_fuzz.py 75 > gpurun_out/r02ai_dist_fuzz.json 2> gpurun_out/r02ai_dist_fuzz.err
echo rc=$?
tail -c 1500 gpurun_out/r02ai_dist_fuzz.json; tail -c 1500 gpurun_out/r02ai_dist_fuzz.err

#!/bin/bash
# round 2: regression test for blank-but-nonempty literals; extended fuzz from seed 14 on
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -q -x -k "blank_but_not_empty or ragged_edge or fixed_strides" 2>&1 | tail -3 )
timeout 500 python tools/gpu_fuzz_more.py 240 14 > gpurun_out/r02af_fuzz.log 2>&1
tail -12 gpurun_out/r02af_fuzz.log

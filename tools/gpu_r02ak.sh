#!/bin/bash
# round 2: ragged batches with long strings mixed in (fuzz); the -m gpu suite on the final library
mkdir -p gpurun_out
timeout 240 python tools/gpu_fuzz_long_strings.py 120 0 > gpurun_out/r02ak_fuzz.log 2>&1
tail -5 gpurun_out/r02ak_fuzz.log
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > gpurun_out/r02ak_pytest.log
cat gpurun_out/r02ak_pytest.log

#!/bin/bash
# round 2: work-budget error path, synccheck in detail
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "work_budget or linear_time or prefix_literal or all_matches or sequential" > gpurun_out/r02o_pytest.log 2>&1
tail -5 gpurun_out/r02o_pytest.log
( timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_smoke.py round2 2>&1 | head -120 ) > gpurun_out/r02o_synccheck.log
head -60 gpurun_out/r02o_synccheck.log

#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 600 python -m pytest tests -m gpu -x -q -k "long_attempts or statemap or c4 or buffer or generated or literal or all_matches" 2>&1 | tail -8 ) > gpurun_out/r02k_pytest.log
B="python bench.py --config c4 --no-cpu --no-e2e --steps 10 --warmup 3"
$B > gpurun_out/r02k_new_c4.json 2> gpurun_out/r02k_new.err
FX_SWEEP_SET2=0 $B > gpurun_out/r02k_new_c4_noset2.json 2>> gpurun_out/r02k_new.err
FX_STATEMAP=0 $B > gpurun_out/r02k_new_c4_nobudget.json 2>> gpurun_out/r02k_new.err
FX_STATEMAP=0 FX_SWEEP_SET2=0 $B > gpurun_out/r02k_new_c4_nobudget_noset2.json 2>> gpurun_out/r02k_new.err
( cd old_r01 && python bench.py --config c4 --no-cpu --no-e2e --steps 10 --warmup 3 ) > gpurun_out/r02k_old_c4.json 2> gpurun_out/r02k_old.err
tail -3 gpurun_out/r02k_pytest.log
for f in new_c4 new_c4_noset2 new_c4_nobudget new_c4_nobudget_noset2 old_c4; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02k_$f.json")); print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms", d.get("matches"))
except Exception as e: print("$f", "ERR", e)
PY
done

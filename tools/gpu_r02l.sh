#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) > gpurun_out/r02l_pytest.log
B="python bench.py --config c4 --no-cpu --no-e2e --steps 10 --warmup 3"
$B > gpurun_out/r02l_new_c4.json 2> gpurun_out/r02l_new.err
FX_STATEMAP=0 $B > gpurun_out/r02l_new_c4_nobudget.json 2>> gpurun_out/r02l_new.err
tail -4 gpurun_out/r02l_pytest.log
for f in new_c4 new_c4_nobudget; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02l_$f.json")); print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms", d.get("matches"))
except Exception as e: print("$f", "ERR", e)
PY
done

#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 600 python -m pytest tests -m gpu -x -q -k "long_attempts or statemap or c4" 2>&1 | tail -8 ) > gpurun_out/r02h_pytest.log
B="python bench.py --no-cpu --no-e2e --steps 10 --warmup 3"
$B --config c4 > gpurun_out/r02h_c4.json 2> gpurun_out/r02h_c4.err
FX_STATEMAP=0 $B --config c4 > gpurun_out/r02h_c4_nobudget.json 2>> gpurun_out/r02h_c4.err
FX_STATEMAP=0 FX_SWEEP_SET2=0 $B --config c4 > gpurun_out/r02h_c4_nobudget_noset2.json 2>> gpurun_out/r02h_c4.err
FX_SWEEP_SET2=0 $B --config c4 > gpurun_out/r02h_c4_noset2.json 2>> gpurun_out/r02h_c4.err
tail -3 gpurun_out/r02h_pytest.log
for f in c4 c4_nobudget c4_nobudget_noset2 c4_noset2; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02h_$f.json")); t=d["config"]["table"]; print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("verified",{}).get("span_equals_construction"))
except Exception as e: print("$f", "ERR", e)
PY
done

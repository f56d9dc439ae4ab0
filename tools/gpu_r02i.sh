#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu --no-e2e --steps 10 --warmup 3"
for fl in 0 1 2 3; do
FX_K4_FLAGS=$fl $B --config c4 > gpurun_out/r02i_c4_f$fl.json 2>> gpurun_out/r02i_c4.err
done
FX_K4_FLAGS=3 FX_STATEMAP=0 $B --config c4 > gpurun_out/r02i_c4_f3_nobudget.json 2>> gpurun_out/r02i_c4.err
for f in c4_f0 c4_f1 c4_f2 c4_f3 c4_f3_nobudget; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02i_$f.json")); print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("verified",{}).get("span_equals_construction"))
except Exception as e: print("$f", "ERR", e)
PY
done

#!/bin/bash
# round 2: K2c sweep-load cache policy; the alternative ragged forms on the general walker's workload
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export PYTHONUNBUFFERED=1
run() { name=$1; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --config c2 > gpurun_out/r02x_c2_$name.json 2> gpurun_out/r02x_c2_$name.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02x_c2_$name.json") if l.startswith("{")][-1]); g=d.get("general_walker") or {}
    print("$name", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("matches"), "| general", round(g.get("value",0),1), g.get("gpu_results_equal_oracle"))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/r02x_c2_$name.err").read()[-800:])
PY
}
run hint1 FX_SPARSE_STREAM_HINT=1
run hint3 FX_SPARSE_STREAM_HINT=3
run hint2 FX_SPARSE_STREAM_HINT=2
run hint4 FX_SPARSE_STREAM_HINT=4
run form2 FX_RAGGED_FORM=2
run form1 FX_RAGGED_FORM=1

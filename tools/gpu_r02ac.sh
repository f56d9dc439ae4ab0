#!/bin/bash
# round 2: K4 -- unit phase from one unaligned word, attempts bytewise again
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export PYTHONUNBUFFERED=1 FX_BENCH_NO_GENERAL=1
run() { name=$1; lines=$2; shift; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --config c4 --lines $lines > gpurun_out/r02ac_$name.json 2> gpurun_out/r02ac_$name.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02ac_$name.json") if l.startswith("{")][-1]); print("$name", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("verified",{}).get("span_equals_construction"))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/r02ac_$name.err").read()[-800:])
PY
}
run c4_8g 8589934592 FX_X=0
run c4_8g_nostarts 8589934592 FX_K4_PHASES=1
run c4_8g_dense 8589934592 FX_SPARSE=0 FX_STATEMAP=0
run c4_32g 34359738368 FX_X=0
( timeout 900 python -m pytest tests -m gpu -q -x -k "c4 or buffer or prefix or window or generated or sparse or statemap" 2>&1 | tail -3 )

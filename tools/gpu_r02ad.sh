#!/bin/bash
# round 2: K4 with the general attempt loop out of line (smaller kernel): A = 78 registers (3 CTAs/SM), B = capped at 64, 0 = as committed
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 FX_BENCH_NO_GENERAL=1
run() { name=$1; lines=$2; shift; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --config c4 --lines $lines > gpurun_out/r02ad_$name.json 2> gpurun_out/r02ad_$name.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02ad_$name.json") if l.startswith("{")][-1]); print("$name", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("verified",{}).get("span_equals_construction"))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/r02ad_$name.err").read()[-800:])
PY
}
for v in 0 A B; do
  cp build/variants/lib$v.so forgex_b200/libforgex_b200.so
  run c4_8g_$v 8589934592 FX_X=0
done

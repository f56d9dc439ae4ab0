#!/bin/bash
# A/B at the same size on the same box: round-1 tree (old_r01/) vs the current tree, C4 at 32 GiB and 8 GiB
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cp MEASURED_PEAKS.json old_r01/ 2>/dev/null
( cd old_r01 && python bench.py --config c4 --no-cpu --no-e2e --steps 10 --warmup 3 ) > gpurun_out/r02j_old_c4.json 2> gpurun_out/r02j_old.err
( cd old_r01 && python bench.py --config c4 --lines 8589934592 --no-cpu --no-e2e --steps 10 --warmup 3 ) > gpurun_out/r02j_old_c4_8g.json 2>> gpurun_out/r02j_old.err
python bench.py --config c4 --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r02j_new_c4.json 2> gpurun_out/r02j_new.err
python bench.py --config c4 --lines 8589934592 --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r02j_new_c4_8g.json 2>> gpurun_out/r02j_new.err
FX_K4_PHASES=1 python bench.py --config c4 --lines 8589934592 --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r02j_new_c4_8g_nostarts.json 2>> gpurun_out/r02j_new.err
FX_K4_PHASES=0 python bench.py --config c4 --lines 8589934592 --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r02j_new_c4_8g_sweep.json 2>> gpurun_out/r02j_new.err
( cd old_r01 && python bench.py --config c4 --no-cpu --no-e2e --steps 10 --warmup 3 ) > gpurun_out/r02j_old_c4_again.json 2>> gpurun_out/r02j_old.err
for f in old_c4 old_c4_8g new_c4 new_c4_8g new_c4_8g_nostarts new_c4_8g_sweep old_c4_again; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02j_$f.json")); print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms", d.get("matches"), d.get("clocks"))
except Exception as e: print("$f", "ERR", e)
PY
done
tail -3 gpurun_out/r02j_old.err

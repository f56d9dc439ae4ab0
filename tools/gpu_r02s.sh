#!/bin/bash
# round 2: K3f -- books from the looked-up words (no second walk), longest-first claim order, round length, 24 warps
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export PYTHONUNBUFFERED=1 FX_BENCH_NO_GENERAL=1
( timeout 900 python -m pytest tests -m gpu -q -x -k "c3 or span or generated or utf8 or ragged_edge or degenerate or statemap or device_pointer or all_matches" 2>&1 | tail -6 ) > gpurun_out/r02s_pytest.log
tail -3 gpurun_out/r02s_pytest.log
B="python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --config c3"
run() { name=$1; shift; env "$@" $B > gpurun_out/r02s_c3_$name.json 2> gpurun_out/r02s_c3_$name.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02s_c3_$name.json") if l.startswith("{")][-1]); print("$name", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("matches"))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/r02s_c3_$name.err").read()[-800:])
PY
}
run default FX_X=0
run round16 FX_SPAN_ROUND=16
run round64 FX_SPAN_ROUND=64
run warps24 FX_SPAN_WARPS=24
run warps24_round16 FX_SPAN_WARPS=24 FX_SPAN_ROUND=16
run spt32 FX_TILE_STRINGS=32
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1
$NCU -k regex:k_span_ragged -s 1 -c 1 -f -o gpurun_out/r02s_prof_c3 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02s_ncu_c3.log 2>&1
python tools/ncu_summary.py gpurun_out/r02s_prof_c3.ncu-rep > gpurun_out/r02s_prof_c3.txt 2>&1
grep -E "time_duration|inst_executed.sum|per_inst_executed|issue_active" gpurun_out/r02s_prof_c3.txt
FX_SPAN_WARPS=24 $NCU -k regex:k_span_ragged -s 1 -c 1 -f -o gpurun_out/r02s_prof_c3_w24 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02s_ncu_c3_w24.log 2>&1
python tools/ncu_summary.py gpurun_out/r02s_prof_c3_w24.ncu-rep > gpurun_out/r02s_prof_c3_w24.txt 2>&1
grep -E "time_duration|inst_executed.sum|per_inst_executed|issue_active" gpurun_out/r02s_prof_c3_w24.txt
rm -f gpurun_out/r02s_prof_c3_w24.ncu-rep
while [ "$(du -sm gpurun_out | cut -f1)" -gt 50 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; rm -f "$big"
done

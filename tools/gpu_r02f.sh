#!/bin/bash
# round 2, GPU call F: K3f streaming ring, K1c compact table, K4 (batch attempts + unit data in queue)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r02f_pytest.log
B="python bench.py --no-cpu --no-e2e --steps 10 --warmup 3"
$B --config c3 > gpurun_out/r02f_c3.json 2> gpurun_out/r02f_c3.err
FX_SPAN_WARPS=32 $B --config c3 > gpurun_out/r02f_c3_w32.json 2>> gpurun_out/r02f_c3.err
FX_SPAN_STREAM=0 $B --config c3 > gpurun_out/r02f_c3_tiles.json 2>> gpurun_out/r02f_c3.err
FX_SPAN_FK=1 $B --config c3 > gpurun_out/r02f_c3_fk1.json 2>> gpurun_out/r02f_c3.err
$B --config c4 > gpurun_out/r02f_c4.json 2> gpurun_out/r02f_c4.err
FX_K4_PHASES=1 $B --config c4 > gpurun_out/r02f_c4_nostarts.json 2>> gpurun_out/r02f_c4.err
$B --config c5 > gpurun_out/r02f_c5.json 2> gpurun_out/r02f_c5.err
FX_COMPACT=0 $B --config c5 > gpurun_out/r02f_c5_full.json 2>> gpurun_out/r02f_c5.err
FX_C5_AUTO=1 $B --config c5 > gpurun_out/r02f_c5_smem.json 2>> gpurun_out/r02f_c5.err
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1
$NCU -k regex:k_span_stream -s 1 -c 1 -f -o gpurun_out/r02f_prof_c3 python bench.py --config c3 --lines 2000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02f_ncu_c3.log 2>&1
$NCU -k regex:k_buffer_scan_sparse -s 2 -c 1 -f -o gpurun_out/r02f_prof_c4 python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02f_ncu_c4.log 2>&1
$NCU -k regex:k_bool_fixed_compact -s 1 -c 1 -f -o gpurun_out/r02f_prof_c5 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02f_ncu_c5.log 2>&1
for c in c3 c4 c5; do python tools/ncu_summary.py gpurun_out/r02f_prof_$c.ncu-rep > gpurun_out/r02f_prof_$c.txt 2>&1; done
while [ "$(du -sm gpurun_out | cut -f1)" -gt 50 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; rm -f "$big"
done
tail -5 gpurun_out/r02f_pytest.log
for f in c3 c3_w32 c3_tiles c3_fk1 c4 c4_nostarts c5 c5_full c5_smem; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02f_$f.json")); t=d["config"]["table"]; print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("verified",{}).get("span_equals_construction"), "compact", t.get("compact_used"), "res", t.get("residency"))
except Exception as e: print("$f", "ERR", e)
PY
done

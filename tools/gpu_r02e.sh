#!/bin/bash
# round 2, GPU call E: parity; K4 v3 without spills; K5 with the single-trajectory phase; K3f pooled vs warp tiles
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r02e_pytest.log
B="python bench.py --no-cpu --no-e2e --steps 10 --warmup 3"
$B --config c4 > gpurun_out/r02e_c4.json 2> gpurun_out/r02e_c4.err
FX_SWEEP_SET2=0 $B --config c4 > gpurun_out/r02e_c4_noset2.json 2>> gpurun_out/r02e_c4.err
FX_K4_PHASES=1 $B --config c4 > gpurun_out/r02e_c4_nostarts.json 2>> gpurun_out/r02e_c4.err
FX_STATEMAP=2 $B --config c4 > gpurun_out/r02e_c4_statemap.json 2>> gpurun_out/r02e_c4.err
$B --config c3 > gpurun_out/r02e_c3.json 2> gpurun_out/r02e_c3.err
FX_SPAN_POOL=0 $B --config c3 > gpurun_out/r02e_c3_nopool.json 2>> gpurun_out/r02e_c3.err
FX_SPAN_FK=1 $B --config c3 > gpurun_out/r02e_c3_fk1.json 2>> gpurun_out/r02e_c3.err
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1
$NCU -k regex:k_buffer_scan_sparse -s 2 -c 1 -f -o gpurun_out/r02e_prof_c4 python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02e_ncu_c4.log 2>&1
FX_STATEMAP=2 $NCU -k regex:k_statemap_regions -s 1 -c 1 -f -o gpurun_out/r02e_prof_c4_statemap python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02e_ncu_c4sm.log 2>&1
$NCU -k regex:k_span_pool -s 1 -c 1 -f -o gpurun_out/r02e_prof_c3 python bench.py --config c3 --lines 2000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02e_ncu_c3.log 2>&1
for c in c4 c4_statemap c3; do python tools/ncu_summary.py gpurun_out/r02e_prof_$c.ncu-rep > gpurun_out/r02e_prof_$c.txt 2>&1; done
rm -f gpurun_out/r02e_prof_c4_statemap.ncu-rep
while [ "$(du -sm gpurun_out | cut -f1)" -gt 50 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; rm -f "$big"
done
tail -5 gpurun_out/r02e_pytest.log
for f in c4 c4_noset2 c4_nostarts c4_statemap c3 c3_nopool c3_fk1; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02e_$f.json")); print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("verified",{}).get("span_equals_construction"), d["config"]["table"].get("statemap_used"))
except Exception as e: print("$f", "ERR", e)
PY
done

#!/bin/bash
# round 2, GPU call B: K3f v2 (span words, lane pool), padded u8 rows, K1 consecutive strings, K4 phase experiments
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r02b_pytest.log
B="python bench.py --no-cpu --no-e2e --steps 10 --warmup 3"
$B --config c3 > gpurun_out/r02b_c3.json 2> gpurun_out/r02b_c3.err
FX_SPAN_FK=1 $B --config c3 > gpurun_out/r02b_c3_fk1.json 2>> gpurun_out/r02b_c3.err
FX_TILE_STRINGS=32 $B --config c3 > gpurun_out/r02b_c3_spt32.json 2>> gpurun_out/r02b_c3.err
$B --config c1 > gpurun_out/r02b_c1.json 2> gpurun_out/r02b_c1.err
FX_SPARSE=0 $B --config c2 --lines 20000000 > gpurun_out/r02b_c2_k2.json 2> gpurun_out/r02b_c2.err
$B --config c4 --lines 8589934592 > gpurun_out/r02b_c4.json 2> gpurun_out/r02b_c4.err
FX_K4_PHASES=1 $B --config c4 --lines 8589934592 > gpurun_out/r02b_c4_nostarts.json 2>> gpurun_out/r02b_c4.err
FX_K4_PHASES=0 $B --config c4 --lines 8589934592 > gpurun_out/r02b_c4_sweeponly.json 2>> gpurun_out/r02b_c4.err
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1
$NCU -k regex:k_span_ragged -s 1 -c 1 -f -o gpurun_out/r02b_prof_c3 python bench.py --config c3 --lines 2000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02b_ncu_c3.log 2>&1
FX_K4_PHASES=1 $NCU -k regex:k_buffer_scan_sparse -s 2 -c 1 -f -o gpurun_out/r02b_prof_c4_nostarts python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02b_ncu_c4a.log 2>&1
FX_K4_PHASES=0 $NCU -k regex:k_buffer_scan_sparse -s 2 -c 1 -f -o gpurun_out/r02b_prof_c4_sweeponly python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02b_ncu_c4b.log 2>&1
$NCU -k regex:k_buffer_scan_sparse -s 2 -c 1 -f -o gpurun_out/r02b_prof_c4 python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02b_ncu_c4.log 2>&1
$NCU -k regex:k_bool_fixed -s 1 -c 1 -f -o gpurun_out/r02b_prof_c1 python bench.py --config c1 --lines 134217728 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02b_ncu_c1.log 2>&1
for c in c3 c4_nostarts c4_sweeponly c4 c1; do python tools/ncu_summary.py gpurun_out/r02b_prof_$c.ncu-rep > gpurun_out/r02b_prof_$c.txt 2>&1; done
# keep the c3 and c4 reports if they fit
rm -f gpurun_out/r02b_prof_c4_nostarts.ncu-rep gpurun_out/r02b_prof_c4_sweeponly.ncu-rep gpurun_out/r02b_prof_c1.ncu-rep
while [ "$(du -sm gpurun_out | cut -f1)" -gt 50 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; rm -f "$big"
done
tail -5 gpurun_out/r02b_pytest.log
for f in c3 c3_fk1 c3_spt32 c1 c2_k2 c4 c4_nostarts c4_sweeponly; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02b_$f.json")); print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("verified",{}).get("span_equals_construction"))
except Exception as e: print("$f", "ERR", e)
PY
done

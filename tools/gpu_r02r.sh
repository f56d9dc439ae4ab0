#!/bin/bash
# round 2: K4 with the L1-allocating sweep load at full size, the same question for K2c, K3f source-level hot spots
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export PYTHONUNBUFFERED=1 FX_BENCH_NO_GENERAL=1
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) > gpurun_out/r02r_pytest.log
tail -3 gpurun_out/r02r_pytest.log
B="python bench.py --no-cpu --no-e2e --steps 10 --warmup 3"
$B --config c4 > gpurun_out/r02r_c4.json 2> gpurun_out/r02r_c4.err
FX_SPARSE_STREAM_HINT=1 $B --config c2 > gpurun_out/r02r_c2_hint1.json 2> gpurun_out/r02r_c2.err
FX_SPARSE_STREAM_HINT=0 $B --config c2 > gpurun_out/r02r_c2_hint0.json 2>> gpurun_out/r02r_c2.err
for f in c4 c2_hint1 c2_hint0; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02r_$f.json") if l.startswith("{")][-1]); print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3))
except Exception as e: print("$f ERR", e)
PY
done
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1
$NCU -k regex:k_span_ragged -s 1 -c 1 -f -o gpurun_out/r02r_prof_c3 python bench.py --config c3 --lines 4000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02r_ncu_c3.log 2>&1
ncu -i gpurun_out/r02r_prof_c3.ncu-rep --page source --csv > gpurun_out/r02r_c3_source.csv 2> gpurun_out/r02r_c3_source.err
FX_SPARSE_STREAM_HINT=0 $NCU -k regex:k_in_sparse -s 1 -c 1 -f -o gpurun_out/r02r_prof_c2_hint0 python bench.py --config c2 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02r_ncu_c2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02r_prof_c2_hint0.ncu-rep > gpurun_out/r02r_prof_c2_hint0.txt 2>&1
grep -E "time_duration|dram__bytes" gpurun_out/r02r_prof_c2_hint0.txt
rm -f gpurun_out/r02r_prof_c2_hint0.ncu-rep
ls -la gpurun_out/*.ncu-rep gpurun_out/r02r_c3_source.csv
while [ "$(du -sm gpurun_out | cut -f1)" -gt 50 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; rm -f "$big"
done

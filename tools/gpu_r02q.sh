#!/bin/bash
# round 2: K4 sweep-load cache policy experiment (the unit phase's re-reads come from DRAM: profiles/traffic.json c4)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export PYTHONUNBUFFERED=1 FX_BENCH_NO_GENERAL=1
B="python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --config c4 --lines 8589934592"
for lp in 0 1 2 3 4 0 1; do
  FX_K4_LOAD=$lp $B > gpurun_out/r02q_c4_lp$lp.json 2> gpurun_out/r02q_c4_lp$lp.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02q_c4_lp$lp.json") if l.startswith("{")][-1]); print("lp$lp", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms", d.get("verified"))
except Exception as e: print("lp$lp ERR", e); print(open("gpurun_out/r02q_c4_lp$lp.err").read()[-1500:])
PY
done
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1
for lp in 0 1 3; do
FX_K4_LOAD=$lp $NCU -k regex:k_buffer_scan_sparse -s 2 -c 1 -f -o gpurun_out/r02q_prof_c4_lp$lp python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02q_ncu_c4_lp$lp.log 2>&1
python tools/ncu_summary.py gpurun_out/r02q_prof_c4_lp$lp.ncu-rep > gpurun_out/r02q_prof_c4_lp$lp.txt 2>&1
echo lp$lp; grep -E "time_duration|dram__bytes_read|hit_rate|lookup_miss|op_read.sum|long_scoreboard" gpurun_out/r02q_prof_c4_lp$lp.txt
done
rm -f gpurun_out/*.ncu-rep

#!/bin/bash
# round 2: K1c global form -- L1 carve-out and evict_last table loads
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export PYTHONUNBUFFERED=1
run() { name=$1; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --config c5 > gpurun_out/r02z_c5_$name.json 2> gpurun_out/r02z_c5_$name.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02z_c5_$name.json") if l.startswith("{")][-1])
    print("$name", round(d["value"],1), "GB/s", round(d["ms_per_step"],4), "ms frac", round(d["roofline"]["frac"],3), d.get("matches"))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/r02z_c5_$name.err").read()[-800:])
PY
}
run maxl1 FX_C5_MAXL1=1
run nomaxl1 FX_C5_MAXL1=0
run smem FX_C5_AUTO=1
( timeout 600 python -m pytest tests -m gpu -q -x -k "c5 or fixed" 2>&1 | tail -3 )
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1 FX_BENCH_NO_GENERAL=1
$NCU -k regex:k_bool_fixed_compact -s 1 -c 1 -f -o gpurun_out/r02z_prof_c5 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02z_ncu_c5.log 2>&1
python tools/ncu_summary.py gpurun_out/r02z_prof_c5.ncu-rep > gpurun_out/r02z_prof_c5.txt 2>&1
grep -E "time_duration|hit_rate|long_scoreboard|issue_active|dram__bytes_read" gpurun_out/r02z_prof_c5.txt
rm -f gpurun_out/*.ncu-rep

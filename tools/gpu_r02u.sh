#!/bin/bash
# round 2: K5 -- block-wise boundary scan, next block in flight, heads/tails from registers, batched compose
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export PYTHONUNBUFFERED=1 FX_BENCH_NO_GENERAL=1
( timeout 900 python -m pytest tests -m gpu -q -x -k "statemap or linear_time or prefix_literal or all_matches or work_budget or c4 or buffer" 2>&1 | tail -6 ) > gpurun_out/r02u_pytest.log
tail -3 gpurun_out/r02u_pytest.log
run() { name=$1; lines=$2; shift; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 --config c4 --lines $lines > gpurun_out/r02u_$name.json 2> gpurun_out/r02u_$name.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02u_$name.json") if l.startswith("{")][-1]); print("$name", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms", d.get("verified",{}).get("span_equals_construction"))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/r02u_$name.err").read()[-800:])
PY
}
run statemap_4g 4294967296 FX_STATEMAP=2
run statemap_32g 34359738368 FX_STATEMAP=2
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1
FX_STATEMAP=2 $NCU -k regex:k_statemap_regions -s 1 -c 1 -f -o gpurun_out/r02u_prof_c4_statemap python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02u_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02u_prof_c4_statemap.ncu-rep > gpurun_out/r02u_prof_c4_statemap.txt 2>&1
grep -E "time_duration|inst_executed.sum|per_inst_executed|issue_active|long_scoreboard" gpurun_out/r02u_prof_c4_statemap.txt
FX_STATEMAP=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02u_launches_statemap.csv python bench.py --config c4 --lines 4294967296 --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
grep -E "statemap|finish" gpurun_out/r02u_launches_statemap.csv | tail -6 | cut -c1-200

#!/bin/bash
# round 2: K5 with the two-level compose
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export PYTHONUNBUFFERED=1 FX_BENCH_NO_GENERAL=1
( timeout 900 python -m pytest tests -m gpu -q -x -k "statemap or linear_time or prefix_literal or all_matches or work_budget or c4 or buffer or generated" 2>&1 | tail -6 ) > gpurun_out/r02v_pytest.log
tail -3 gpurun_out/r02v_pytest.log
run() { name=$1; lines=$2; shift; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 --config c4 --lines $lines > gpurun_out/r02v_$name.json 2> gpurun_out/r02v_$name.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02v_$name.json") if l.startswith("{")][-1]); print("$name", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms", d.get("verified",{}).get("span_equals_construction"))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/r02v_$name.err").read()[-800:])
PY
}
run statemap_64m 67108864 FX_STATEMAP=2
run statemap_4g 4294967296 FX_STATEMAP=2
run statemap_32g 34359738368 FX_STATEMAP=2
FX_STATEMAP=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02v_launches_statemap.csv python bench.py --config c4 --lines 4294967296 --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02v_launches_statemap.csv")))
for r in rows:
    if len(r)>5 and ("statemap" in r[4] or "finish" in r[4]):
        print(r[4][:40], r[-1])
PY

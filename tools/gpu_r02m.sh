#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) > gpurun_out/r02m_pytest.log
B="python bench.py --config c4 --no-cpu --no-e2e --steps 10 --warmup 3"
FX_STATEMAP=2 $B > gpurun_out/r02m_c4_statemap.json 2> gpurun_out/r02m.err
$B > gpurun_out/r02m_c4.json 2>> gpurun_out/r02m.err
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1
FX_STATEMAP=2 $NCU -k regex:k_statemap_regions -s 1 -c 1 -f -o gpurun_out/r02m_prof_c4_statemap python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02m_ncu_c4sm.log 2>&1
python tools/ncu_summary.py gpurun_out/r02m_prof_c4_statemap.ncu-rep > gpurun_out/r02m_prof_c4_statemap.txt 2>&1
rm -f gpurun_out/*.ncu-rep
tail -4 gpurun_out/r02m_pytest.log
for f in c4_statemap c4; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02m_$f.json")); print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms", d.get("matches"), d.get("verified",{}).get("span_equals_construction"))
except Exception as e: print("$f", "ERR", e)
PY
done
grep -E "gpu__time_duration|smsp__inst_executed.sum|thread_inst_executed_per|issue_active" gpurun_out/r02m_prof_c4_statemap.txt | cut -c1-130

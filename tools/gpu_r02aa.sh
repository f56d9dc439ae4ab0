#!/bin/bash
# round 2: K4 -- attempts' plain stretch and the unit phase's four steps from registers
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export PYTHONUNBUFFERED=1 FX_BENCH_NO_GENERAL=1
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) > gpurun_out/r02aa_pytest.log
tail -3 gpurun_out/r02aa_pytest.log
run() { name=$1; lines=$2; shift; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --config c4 --lines $lines > gpurun_out/r02aa_$name.json 2> gpurun_out/r02aa_$name.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02aa_$name.json") if l.startswith("{")][-1]); print("$name", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("verified",{}).get("span_equals_construction"))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/r02aa_$name.err").read()[-800:])
PY
}
run c4_8g 8589934592 FX_X=0
run c4_32g 34359738368 FX_X=0
run c4_8g_nostarts 8589934592 FX_K4_PHASES=1
run c4_8g_dense 8589934592 FX_SPARSE=0 FX_STATEMAP=0
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1
$NCU -k regex:k_buffer_scan_sparse -s 2 -c 1 -f -o gpurun_out/r02aa_prof_c4 python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02aa_ncu_c4.log 2>&1
python tools/ncu_summary.py gpurun_out/r02aa_prof_c4.ncu-rep > gpurun_out/r02aa_prof_c4.txt 2>&1
grep -E "time_duration|inst_executed.sum|per_inst_executed|issue_active|long_scoreboard|dram__bytes_read" gpurun_out/r02aa_prof_c4.txt
rm -f gpurun_out/*.ncu-rep

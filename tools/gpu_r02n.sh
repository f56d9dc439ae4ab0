#!/bin/bash
# round 2, final single-GPU call: parity suite, the default bench line, launch list, full-size ncu captures (traffic), sanitizer
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r02n_pytest.log
python bench.py > gpurun_out/r02n_bench_all.json 2> gpurun_out/r02n_bench_all.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02n_bench_reference.json 2> gpurun_out/r02n_bench_reference.err
FX_C5_AUTO=1 python bench.py --config c5 --no-cpu > gpurun_out/r02n_bench_c5_smem.json 2> gpurun_out/r02n_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02n_launches_default_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02n_bench_under_ncu.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1 FX_BENCH_NO_GENERAL=1
$NCU -k regex:k_bool_fixed -s 1 -c 1 -f -o gpurun_out/r02n_prof_c1 python bench.py --config c1 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02n_ncu_c1.log 2>&1
$NCU -k regex:k_in_sparse -s 1 -c 1 -f -o gpurun_out/r02n_prof_c2 python bench.py --config c2 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02n_ncu_c2.log 2>&1
$NCU -k regex:k_span_ragged -s 1 -c 1 -f -o gpurun_out/r02n_prof_c3 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02n_ncu_c3.log 2>&1
$NCU -k regex:k_buffer_scan_sparse -s 2 -c 1 -f -o gpurun_out/r02n_prof_c4 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02n_ncu_c4.log 2>&1
$NCU -k regex:k_bool_fixed_compact -s 1 -c 1 -f -o gpurun_out/r02n_prof_c5 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02n_ncu_c5.log 2>&1
unset FX_BENCH_NO_GENERAL
FX_SPARSE=0 $NCU -k regex:k_bool_ragged -s 1 -c 1 -f -o gpurun_out/r02n_prof_c2_k2 python bench.py --config c2 --lines 20000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02n_ncu_c2k2.log 2>&1
for c in c1 c2 c3 c4 c5 c2_k2; do python tools/ncu_summary.py gpurun_out/r02n_prof_$c.ncu-rep > gpurun_out/r02n_prof_$c.txt 2>&1; done
rm -f gpurun_out/*.ncu-rep
( compute-sanitizer --tool memcheck python tools/sanitize_smoke.py round2 2>&1 | tail -6 ) > gpurun_out/r02n_sanitizer_memcheck.log
( compute-sanitizer --tool racecheck python tools/sanitize_smoke.py round2 2>&1 | tail -6 ) > gpurun_out/r02n_sanitizer_racecheck.log
( compute-sanitizer --tool synccheck --num-cuda-barriers 65536 python tools/sanitize_smoke.py round2 2>&1 | tail -6 ) > gpurun_out/r02n_sanitizer_synccheck.log
python tools/sanitize_smoke.py round2 > gpurun_out/r02n_sanitizer_native.log 2>&1
tail -3 gpurun_out/r02n_pytest.log
tail -2 gpurun_out/r02n_sanitizer_*.log
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02n_bench_all.json") if l.startswith("{")][-1])
    for c,r in d["per_config"].items():
        print(c, round(r["value"],1), "GB/s", round(r["ms_per_step"],3), "ms frac", round(r["roofline"]["frac"],3), "e2e", round(r["e2e"]["value"],1), "cpu", round(r["cpu_baseline"]["value"],4), r["cpu_baseline"].get("gpu_results_equal_oracle"))
except Exception as e:
    print("bench ERR", e); print(open("gpurun_out/r02n_bench_all.err").read()[-2000:])
PY
for c in c1 c2 c3 c4 c5 c2_k2; do echo $c; grep -E "kernel:|gpu__time_duration|dram__bytes" gpurun_out/r02n_prof_$c.txt | cut -c1-120; done

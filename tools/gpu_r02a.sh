#!/bin/bash
# round 2, GPU call A: parity tests, the default bench line (all five configs), fresh ncu captures of the current kernels
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r02a_pytest.log
( timeout 900 python bench.py 2> gpurun_out/r02a_bench_all.err ) > gpurun_out/r02a_bench_all.json
NCU="ncu --set full --clock-control none --import-source on"
export FX_BENCH_ALLOW_SHORT_WARMUP=1
$NCU -k regex:k_bool_fixed -s 1 -c 1 -f -o gpurun_out/r02a_prof_c1 python bench.py --config c1 --lines 134217728 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02a_ncu_c1.log 2>&1
$NCU -k regex:k_span_ragged -s 1 -c 1 -f -o gpurun_out/r02a_prof_c3 python bench.py --config c3 --lines 2000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02a_ncu_c3.log 2>&1
$NCU -k regex:k_buffer_scan_sparse -s 1 -c 1 -f -o gpurun_out/r02a_prof_c4 python bench.py --config c4 --lines 2147483648 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02a_ncu_c4.log 2>&1
$NCU -k regex:k_bool_fixed -s 1 -c 1 -f -o gpurun_out/r02a_prof_c5 python bench.py --config c5 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02a_ncu_c5.log 2>&1
FX_SPARSE=0 $NCU -k regex:k_bool_ragged -s 1 -c 1 -f -o gpurun_out/r02a_prof_c2_k2 python bench.py --config c2 --lines 4000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02a_ncu_c2k2.log 2>&1
for c in c1 c3 c4 c5 c2_k2; do python tools/ncu_summary.py gpurun_out/r02a_prof_$c.ncu-rep > gpurun_out/r02a_prof_$c.txt 2>&1; done
tail -5 gpurun_out/r02a_pytest.log
head -c 600 gpurun_out/r02a_bench_all.json
# gpurun_out must stay under 64 MiB: the summaries matter, the reports only as far as they fit
while [ "$(du -sm gpurun_out | cut -f1)" -gt 50 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break; rm -f "$big"
done

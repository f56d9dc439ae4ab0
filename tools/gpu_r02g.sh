#!/bin/bash
# round 2, GPU call G: parity, the default bench line (all five configs + the general walker), K4 with the coarse budget tick
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r02g_pytest.log
( time python bench.py > gpurun_out/r02g_bench_all.json 2> gpurun_out/r02g_bench_all.err ) 2> gpurun_out/r02g_bench_time.log
B="python bench.py --no-cpu --no-e2e --steps 10 --warmup 3"
FX_K4_PHASES=1 $B --config c4 > gpurun_out/r02g_c4_nostarts.json 2> gpurun_out/r02g_c4.err
FX_C5_AUTO=1 $B --config c5 > gpurun_out/r02g_c5_smem.json 2> gpurun_out/r02g_c5.err
FX_L2_PERSIST=0 $B --config c5 > gpurun_out/r02g_c5_nopersist.json 2>> gpurun_out/r02g_c5.err
FX_COMPACT=0 $B --config c5 > gpurun_out/r02g_c5_full.json 2>> gpurun_out/r02g_c5.err
FX_COMPACT=0 FX_L2_PERSIST=0 $B --config c5 > gpurun_out/r02g_c5_full_nopersist.json 2>> gpurun_out/r02g_c5.err
tail -3 gpurun_out/r02g_pytest.log; cat gpurun_out/r02g_bench_time.log | tail -4
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02g_bench_all.json") if l.startswith("{")][-1])
    for c,r in d["per_config"].items():
        print(c, round(r["value"],1), "GB/s", round(r["ms_per_step"],3), "ms frac", round(r["roofline"]["frac"],3), "e2e", round(r["e2e"]["value"],1), "cpu", round(r["cpu_baseline"]["value"],4), r["cpu_baseline"].get("gpu_results_equal_oracle"), "wall", round(r["wall_seconds_incl_setup"],1))
        if "general_walker" in r: print("   general", r["general_walker"])
        if "verified" in r: print("   verified", {k:v for k,v in r["verified"].items() if k!="oracle_slice"}, r["verified"].get("oracle_slice",{}).get("equal"))
except Exception as e:
    print("bench ERR", e); print(open("gpurun_out/r02g_bench_all.err").read()[-2000:])
PY
for f in c4_nostarts c5_smem c5_nopersist c5_full c5_full_nopersist; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02g_$f.json")); t=d["config"]["table"]; print("$f", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), "compact", t.get("compact_used"), "res", t.get("residency"))
except Exception as e: print("$f", "ERR", e)
PY
done

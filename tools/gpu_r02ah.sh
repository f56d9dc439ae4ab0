#!/bin/bash
# round 2: extended fuzz of the fixed-stride kernels; quick bench sanity after the generic_mode fix
mkdir -p gpurun_out
timeout 300 python tools/gpu_fuzz_fixed.py 150 0 > gpurun_out/r02ah_fuzz.log 2>&1
tail -6 gpurun_out/r02ah_fuzz.log
python bench.py --config c1 --no-cpu --no-e2e > gpurun_out/r02ah_c1.json 2> gpurun_out/r02ah_c1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02ah_c1.json") if l.startswith("{")][-1]); print("c1", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3))
PY

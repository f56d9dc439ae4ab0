#!/bin/bash
# round 2: K2c launch knobs on C2 (tile size per warp, queue flush threshold)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 FX_BENCH_NO_GENERAL=1
run() { name=$1; shift; env "$@" python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 --config c2 > gpurun_out/r02al_c2_$name.json 2> gpurun_out/r02al_c2_$name.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02al_c2_$name.json") if l.startswith("{")][-1])
    print("$name", round(d["value"],1), "GB/s", round(d["ms_per_step"],3), "ms frac", round(d["roofline"]["frac"],3), d.get("matches"))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/r02al_c2_$name.err").read()[-500:])
PY
}
run default FX_X=0
run tile12k FX_SPARSE_TILE_BYTES=12288
run tile48k FX_SPARSE_TILE_BYTES=49152
run tile96k FX_SPARSE_TILE_BYTES=98304
run flush16 FX_SPARSE_FLUSH=16
run flush48 FX_SPARSE_FLUSH=48
run tile48k_flush32 FX_SPARSE_TILE_BYTES=49152 FX_SPARSE_FLUSH=32

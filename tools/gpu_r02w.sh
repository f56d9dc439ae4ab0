#!/bin/bash
# round 2: the default bench line on 4 GPUs exactly as the driver launches it (c4 = ONE 32 GiB text in 4 slabs)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export PYTHONUNBUFFERED=1
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02w_bench_${N}gpus.json 2> gpurun_out/r02w_bench_${N}gpus.err
echo rc=$?
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02w_bench_${N}gpus.json") if l.startswith("{")][-1])
    print("headline", round(d["value"],1), d["n_gpus"], "e2e", round(d["e2e"]["value"],1))
    for c,r in d["per_config"].items():
        print(c, round(r["value"],1), "GB/s", round(r["ms_per_step"],3), "ms", r["scaling"], "e2e", round(r["e2e"]["value"],1), r.get("verified",{}).get("span_equals_construction"), r.get("collectives"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/r02w_bench_${N}gpus.err").read()[-3000:])
PY
tail -c 600 gpurun_out/r02w_bench_${N}gpus.err

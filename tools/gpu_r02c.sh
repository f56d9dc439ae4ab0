#!/bin/bash
# round 2, GPU call C (2 GPUs): the default bench line under torchrun (c4 split over the ranks), the split-buffer demo
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1
./tests/c/cabi_smoke > gpurun_out/r02c_cabi.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02c_bench_2gpu.json 2> gpurun_out/r02c_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_buffer_demo.py > gpurun_out/r02c_dist_demo.json 2> gpurun_out/r02c_dist_demo.err
tail -3 gpurun_out/r02c_smoke.log; cat gpurun_out/r02c_cabi.log
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r02c_bench_2gpu.json"))
    for c,r in d["per_config"].items():
        print(c, round(r["value"],1), "GB/s", round(r["ms_per_step"],3), "ms", r["scaling"], "e2e", round(r["e2e"]["value"],1) if r.get("e2e") else None, r.get("verified"), r.get("collectives"))
    print(d.get("numa_pin"))
except Exception as e:
    print("bench ERR", e); print(open("gpurun_out/r02c_bench_2gpu.err").read()[-3000:])
try:
    print(open("gpurun_out/r02c_dist_demo.json").read()[:1500])
except Exception as e: print(e)
PY
tail -5 gpurun_out/r02c_dist_demo.err

#!/bin/bash
# round 2: synccheck with a barrier table large enough for 148 CTAs x 32 warp-private mbarriers (K3f)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
( timeout 900 compute-sanitizer --tool synccheck --num-cuda-barriers 65536 python tools/sanitize_smoke.py round2 2>&1 | head -120 ) > gpurun_out/r02p_synccheck.log
head -60 gpurun_out/r02p_synccheck.log

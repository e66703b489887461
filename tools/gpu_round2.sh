#!/bin/bash
mkdir -p gpurun_out
for r in 1 4 8 12 16 20 24 32; do TOR_BVH_REFILL=$r python tools/sweep.py; done 2>&1 | tee gpurun_out/sweep_refill.txt
ncu --set full --clock-control none --import-source on -k regex:render_bvh -s 1 -c 1 -f -o gpurun_out/prof_bvh_c2 \
    python tools/sweep.py > gpurun_out/ncu_bvh.log 2>&1
tail -3 gpurun_out/ncu_bvh.log

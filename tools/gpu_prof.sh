#!/bin/bash
# ncu capture of the BVH kernel on a reduced C2 (same scene/camera, 800x450, 32 spp) so that the
# instrumented passes stay short
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:render_bvh -s 1 -c 1 -f -o gpurun_out/prof_bvh_small \
    python tools/sweep.py --dims 450 800 32 2 > gpurun_out/ncu_bvh_small.log 2>&1
tail -3 gpurun_out/ncu_bvh_small.log

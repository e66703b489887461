#!/bin/bash
TAG=${1:-r02e}
mkdir -p gpurun_out
TOR_BVH_PREPASS_SPP=9 TOR_BVH_COOP_FORCE=1000000000 TOR_BVH_COOP_MODE=0 timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:render_bvh -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_coop python tools/sweep.py --dims 216 384 32 1 > gpurun_out/${TAG}_ncu_coop.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_coop.log

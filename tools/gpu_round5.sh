#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_quick.py 2>&1 | cut -c1-120
{
python tools/sweep.py --dims 675 1200 500 1
for cfg in "1 1 1 1 4" "1 1 1 1 1" "4 4 4 4 4" "8 4 8 8 4" "8 4 12 12 8" "8 8 16 16 8" "4 4 16 16 8" "1 4 16 16 4" "1 8 20 20 4" "1 1 16 16 4" "1 1 24 24 8" "1 1 8 8 2"; do
set -- $cfg
TOR_BVH_THR_NODE=$1 TOR_BVH_THR_LEAF=$2 TOR_BVH_THR_HIT=$3 TOR_BVH_THR_NEW=$4 TOR_BVH_NODE_BURST=$5 python tools/sweep.py --dims 675 1200 500 1
done
} 2>&1 | tee gpurun_out/sweep_thr2.txt

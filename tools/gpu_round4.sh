#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_quick.py 2>&1 | cut -c1-200
{
python tools/sweep.py
for n in 8 16 24; do for l in 4 8 12 16 24; do TOR_BVH_THR_NODE=$n TOR_BVH_THR_LEAF=$l python tools/sweep.py --dims 675 1200 500 1; done; done
for h in 4 8 16 24; do TOR_BVH_THR_HIT=$h python tools/sweep.py --dims 675 1200 500 1; done
for w in 2 4 16 24; do TOR_BVH_THR_NEW=$w python tools/sweep.py --dims 675 1200 500 1; done
for b in 1 2 8 16; do TOR_BVH_NODE_BURST=$b python tools/sweep.py --dims 675 1200 500 1; done
} 2>&1 | tee gpurun_out/sweep_thr.txt

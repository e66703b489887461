#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_quick.py 2>&1 | cut -c1-260
for r in 8 12 16 24 32; do TOR_BVH_REFILL=$r python tools/sweep.py; done 2>&1 | tee gpurun_out/sweep_refill2.txt

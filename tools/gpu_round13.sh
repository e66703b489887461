#!/bin/bash
python tools/gpu_quick.py 2>&1 | cut -c1-330
for r in 16 20 24; do TOR_BVH_REFILL=$r python tools/sweep.py --dims 675 1200 500 2; done

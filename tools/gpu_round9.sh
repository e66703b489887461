#!/bin/bash
for L in 32 16 8 4; do for f in "" "--rowmajor"; do TOR_BVH_LANES=$L TOR_BVH_REFILL=$((L*5/8)) python tools/sweep.py --dims 675 1200 500 1 --rowstep 8 $f; done; done
for L in 32 16 8; do for f in "" "--rowmajor"; do TOR_BVH_LANES=$L TOR_BVH_REFILL=$((L*5/8)) python tools/sweep.py --dims 675 1200 500 1 --rowstep 4 $f; done; done
for L in 24 16; do TOR_BVH_LANES=$L TOR_BVH_REFILL=$((L*5/8)) python tools/sweep.py --dims 675 1200 500 1; done
for L in 24 16; do TOR_BVH_LANES=$L TOR_BVH_REFILL=$((L*5/8)) python tools/sweep.py --dims 675 1200 500 1 --rowstep 2; done

#!/bin/bash
for blk in 256 512; do echo "BLOCK $blk"; TOR_BVH_BLOCK=$blk python tools/sweep.py --dims 675 1200 500 3 | tail -1 | cut -c1-200; TOR_BVH_BLOCK=$blk python tools/sweep.py --dims 675 1200 500 3 --fast | tail -1 | cut -c1-200; TOR_BVH_BLOCK=$blk python tools/sweep.py --c1 | tail -1 | cut -c1-200; done

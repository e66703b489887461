#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for k in 1 2 4 8; do python tools/sweep.py --dims 675 1200 500 2 --rowstep $k; done
python tools/sweep.py --c1
TOR_BVH_LANES=32 python tools/sweep.py --c1

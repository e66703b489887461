#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_quick.py 2>&1 | cut -c1-250
for k in 1 2 4 8; do python tools/sweep.py --dims 675 1200 500 2 --rowstep $k; done
python tools/sweep.py --dims 675 1200 500 2 --rowmajor
python tools/sweep.py --c1
python tools/sweep.py --c1 --rowmajor

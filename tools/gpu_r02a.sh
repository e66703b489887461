#!/bin/bash
# Round-2 GPU session A: parity (incl. the warp-cooperative route), then the policy sweep of tools/measure_coop.py.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
timeout 600 python tools/measure_coop.py ${TAG} 2>&1 | tail -80 | tee gpurun_out/${TAG}_coop.log

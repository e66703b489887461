#!/bin/bash
# quick loop: coop parity tests + solo latency + the step-8/4/1 timings for the default policy
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cooperative or cost_ranked or small_random" 2>&1 | tail -3
timeout 300 python tools/measure_latency.py 2>&1 | tail -3
timeout 600 python tools/measure_coop.py ${TAG} --quick 2>&1 | grep -v "^c1" | tail -12

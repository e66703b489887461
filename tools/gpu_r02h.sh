#!/bin/bash
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-seconds 5 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; tail -c 4000 gpurun_out/${TAG}_bench_c2.json; tail -3 gpurun_out/${TAG}_bench_c2.err
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; tail -c 2500 gpurun_out/${TAG}_bench_c4.json; tail -5 gpurun_out/${TAG}_bench_c4.err

#!/bin/bash
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_c2.json"))
print("c2 value", d["value"], "ms", d["ms_per_step"], "split", d["split_stream_mode"]["value"], d["split_stream_mode"]["ms_per_step"], d["image_check"]["result"])
PY
tail -3 gpurun_out/${TAG}_bench_c2.err
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_c4.json 2> gpurun_out/${TAG}_bench_c4.err; tail -c 2500 gpurun_out/${TAG}_bench_c4.json; tail -5 gpurun_out/${TAG}_bench_c4.err
timeout 600 python tools/measure_coop.py ${TAG} --quick 2>&1 | grep -v "^c1" | tail -9

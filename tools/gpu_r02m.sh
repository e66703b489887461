#!/bin/bash
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 900 python tools/sweep_objects.py ${TAG} 2>&1 | tail -12 | tee gpurun_out/${TAG}_object_sweep.md
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err; python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_c5.json"))
print("c5 value", round(d["value"]), "ms", round(d["ms_per_step"]), "split", round(d["split_stream_mode"]["value"]), d["image_check"], d["roofline"]["work"])
PY
tail -3 gpurun_out/${TAG}_bench_c5.err
# FP64 instruction counts + DRAM traffic of one whole C2 render (pre-pass + main launch), exact and split-stream
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_bvh -c 2 -f -o gpurun_out/${TAG}_prof_c2_exact python tools/sweep.py --dims 675 1200 500 1 > gpurun_out/${TAG}_ncu_c2_exact.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_c2_exact.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_bvh -c 1 -f -o gpurun_out/${TAG}_prof_c2_split python tools/sweep.py --dims 675 1200 500 1 --fast > gpurun_out/${TAG}_ncu_c2_split.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_c2_split.log

#!/bin/bash
# Final evidence for the round: parity tests, smoke, bench + reference arm, launch list, full ncu capture.
# usage: tools/gpu_final.sh <tag>   (artefacts land in gpurun_out/<tag>_*)
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 3500 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>&1; tail -c 300 gpurun_out/${TAG}_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
grep -c render_bvh gpurun_out/${TAG}_launches_bench_c2.csv
ncu --set full --clock-control none --import-source on -k regex:render_bvh -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bvh \
    python tools/sweep.py --dims 450 800 128 2 --rowmajor > gpurun_out/${TAG}_ncu_bvh.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bvh.log
# the same capture for the split-stream (chunked queue) variant of the kernel
ncu --set full --clock-control none --import-source on -k regex:render_bvh -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bvh_split \
    python tools/sweep.py --dims 450 800 128 2 --fast > gpurun_out/${TAG}_ncu_bvh_split.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_bvh_split.log
# DRAM traffic of one full C2 launch of the main render kernel (the second render_bvh launch; the first is the cost pre-pass)
ncu --set full --clock-control none -k regex:render_bvh -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_c2_main \
    python tools/sweep.py --dims 675 1200 500 1 > gpurun_out/${TAG}_ncu_c2.log 2>&1
python tools/ncu_traffic.py gpurun_out/${TAG}_prof_c2_main.ncu-rep > gpurun_out/${TAG}_ncu_traffic.json; cat gpurun_out/${TAG}_ncu_traffic.json

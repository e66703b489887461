#!/bin/bash
# Final evidence for the round: parity tests, smoke, bench + reference arm, launch list, full ncu capture.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 300 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c render_bvh gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k regex:render_bvh -s 1 -c 1 -f -o gpurun_out/prof_bvh_r01f \
    python tools/sweep.py --dims 450 800 128 2 > gpurun_out/ncu_bvh_r01f.log 2>&1
tail -2 gpurun_out/ncu_bvh_r01f.log

#!/bin/bash
# One GPU session: parity tests, smoke, bench (N=1) + reference arm, launch list, one full ncu capture.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
python bench.py --route brute --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_brute.json 2>&1; tail -c 1500 gpurun_out/bench_n1_brute.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -25 gpurun_out/launches.csv | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:render_bvh -s 1 -c 1 -f -o gpurun_out/prof_bvh_r01c \
    python tools/sweep.py --dims 450 800 160 2 > gpurun_out/ncu_bvh_r01c.log 2>&1
tail -3 gpurun_out/ncu_bvh_r01c.log

#!/bin/bash
# Final evidence for a round on ONE GPU: parity tests, smoke, bench (C2) + reference arm, launch list, full ncu captures
# of the real C2 launches (exact: pre-pass + main; split-stream), FP64 op counts, DRAM traffic.
# usage: tools/gpu_final.sh <tag>   (artefacts land in gpurun_out/<tag>_*)
TAG=${1:-r02}
REF=${2:-noref}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 3000 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
if [ "$REF" = "ref" ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>&1; tail -c 600 gpurun_out/${TAG}_bench_reference_arm.json
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/${TAG}_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
grep -c render_bvh gpurun_out/${TAG}_launches_bench_c2.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_bvh -c 2 -f -o gpurun_out/${TAG}_prof_c2_exact \
    python tools/sweep.py --dims 675 1200 500 1 > gpurun_out/${TAG}_ncu_c2_exact.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_c2_exact.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_bvh -c 1 -f -o gpurun_out/${TAG}_prof_c2_split \
    python tools/sweep.py --dims 675 1200 500 1 --fast > gpurun_out/${TAG}_ncu_c2_split.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_c2_split.log
# the cooperative kernel + lane kernel of one GPU's share of C2 on 8 GPUs (every 8th row)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_coop -c 1 -f -o gpurun_out/${TAG}_prof_c2_share8_coop \
    python tools/sweep.py --dims 675 1200 500 1 --rowstep 8 > gpurun_out/${TAG}_ncu_c2_share8.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_c2_share8.log

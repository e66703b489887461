#!/bin/bash
# N-GPU bench lines the way the driver launches them (+ the in-process multi-device test)
# usage: tools/gpu_multi.sh N TAG [c5|c2only]
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device" 2>&1 | tail -2
run() {  # workload, port, extra args
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 \
      bench.py --gpus $N --workload $1 ${@:3} > gpurun_out/${TAG}_bench_$1_n$N.json 2> gpurun_out/${TAG}_bench_$1_n$N.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$1_n$N.json"))
    s=d.get("split_stream_mode") or {}
    print("$1 N=$N value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "split", round(s.get("value",0)), "split e2e", round(s.get("e2e",0) or 0), "image", (d.get("image_check") or {}).get("result"), (d.get("image_check") or {}).get("gathered_rows_equal_single_rank_render"), d.get("schedule"), "frac", d["roofline"].get("frac"))
except Exception as e:
    print("$1 N=$N FAILED", e)
PY
  tail -2 gpurun_out/${TAG}_bench_$1_n$N.err | cut -c1-300
}
run c2 29511 --steps 5 --warmup 3
if [ "$3" != "c2only" ]; then run c4 29513 --steps 2 --warmup 1; fi
if [ "$3" = "c5" ]; then run c5 29515 --steps 1 --warmup 1; fi

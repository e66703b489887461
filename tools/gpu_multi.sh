#!/bin/bash
# N-GPU bench lines the way the driver launches them (+ the in-process multi-device test)
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_device" 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n$N.json"))
print("c2 N=$N value", round(d["value"]), "ms", round(d["ms_per_step"],2), "kernel_ms", round(d["roofline"]["kernel_ms"],2), "e2e", round(d["e2e"]["value"]), "split", round(d["split_stream_mode"]["value"]), "split e2e", round(d["split_stream_mode"]["e2e"]), d["image_check"], d.get("schedule"))
PY
tail -3 gpurun_out/${TAG}_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --workload c4 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_c4_n$N.json 2> gpurun_out/${TAG}_bench_c4_n$N.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_c4_n$N.json"))
print("c4 N=$N value", round(d["value"]), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "split", round(d["split_stream_mode"]["value"]), d["image_check"].get("result"))
PY
tail -3 gpurun_out/${TAG}_bench_c4_n$N.err

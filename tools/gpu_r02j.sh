#!/bin/bash
mkdir -p gpurun_out
for st in 2 1 0; do echo "STAGE $st"; TOR_BVH_STAGE=$st python tools/sweep.py --dims 675 1200 500 3 | tail -1; TOR_BVH_STAGE=$st python tools/sweep.py --dims 675 1200 500 3 --fast | tail -1; done
for nf in 2 4 8 16; do echo "in_flight $nf"; python bench.py --workload c4 --steps 1 --warmup 1 --in-flight $nf 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('exact ms', d['ms_per_step'], 'split ms', d['split_stream_mode']['ms_per_step'], d['image_check']['rgb8_sha256_all_frames'][:12])"; done

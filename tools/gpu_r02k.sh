#!/bin/bash
mkdir -p gpurun_out
for cfg in "4 0" "8 0" "8 2" "8 8" "16 0" "16 4" "16 16" "32 0" "32 8"; do set -- $cfg; echo "in_flight $1 div $2"; TOR_ANIM_GRID_DIV=$2 python bench.py --workload c4 --steps 1 --warmup 1 --in-flight $1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('exact ms', d['ms_per_step'], 'split ms', d['split_stream_mode']['ms_per_step'], d['image_check']['rgb8_sha256_all_frames'][:12])"; done

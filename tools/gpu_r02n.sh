#!/bin/bash
for st in 2 1 0; do for half in 22 50; do echo "STAGE<=$st half $half"; TOR_BVH_STAGE=$st python tools/sweep.py --dims 675 1200 100 3 --half $half | tail -1; TOR_BVH_STAGE=$st python tools/sweep.py --dims 675 1200 100 3 --half $half --fast | tail -1; done; done
for st in 1 0; do echo "c4 STAGE<=$st"; TOR_BVH_STAGE=$st python bench.py --workload c4 --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('exact ms', d['ms_per_step'], 'split ms', d['split_stream_mode']['ms_per_step'], d['image_check']['result'])"; done

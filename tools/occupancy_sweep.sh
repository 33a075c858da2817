#!/bin/bash
# occupancy sensitivity of the fused realign kernel: cap resident warps per SM and time the default bench
for w in 6 12 18; do
  INDELGPU_MAX_WARPS_PER_SM=$w timeout 200 python bench.py --steps 5 --warmup 2 --no-cpu 2>/dev/null > /tmp/occ_$w.json
  python -c "import json; d=json.load(open('/tmp/occ_$w.json')); print('warps/SM cap', $w, round(d['ms_per_step'],2), 'ms')"
done

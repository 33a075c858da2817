#!/bin/bash
# e2e sensitivity to the chunk size of indelgpu_realign_batch's overlapped host path
for c in 32768 65536 131072 262144; do
  INDELGPU_CHUNK_READS=$c timeout 200 python bench.py --steps 10 --warmup 2 --no-cpu 2>/dev/null > /tmp/chunk_$c.json
  python -c "import json; d=json.load(open('/tmp/chunk_$c.json')); print('chunk', $c, 'e2e', round(d['e2e']['value']/1e6,1), 'M reads/s; device', round(d['value']/1e6,1))"
done

#!/bin/bash
# e2e (C-ABI host call) per chunk size of the H2D/compute pipeline
for mb in 4 16 64; do
  FO_HOST_CHUNK_MB=$mb python bench.py --no-cpu-baseline --no-stages --no-latency --steps 5 2>/dev/null > /tmp/e2e_$mb.json
  python -c "
import json; d=json.load(open('/tmp/e2e_$mb.json')); print($mb, round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2))"
done

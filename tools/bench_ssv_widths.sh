#!/bin/bash
# SSV tile-width experiment: same workload, CKM_SSV_J = auto / 16 / 32
for J in auto 16 32; do
  echo "== CKM_SSV_J=$J"
  CKM_SSV_J=$J timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'ms/step', round(d['ms_per_step']), 'ssv ms', round(d['gcups']['stage_ms_per_step']['ssv'],1), 'real TCUPS', round(d['gcups']['ssv_gcups_real']/1000,2), 'tile TCUPS', round(d['gcups']['ssv_gcups_tile']/1000,2), d['gcups']['stage_ms_per_step']['last_step'], 'cand', d['cascade']['ssv_cand'])"
done

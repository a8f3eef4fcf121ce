#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t3.log
( time python bench.py --steps 3 --warmup 2 ) > gpurun_out/r2_b3.log 2> gpurun_out/r2_b3.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_b3.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_b3.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
    print('stages',d['gcups']['stage_ms_per_step'])
    print('plugin',d.get('plugin'))
    print('cpu',d.get('cpu_baseline'))
    print('cascade',d['cascade'])
PY
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_b3_ref.log 2>&1; tail -c 1500 gpurun_out/r2_b3_ref.log
( time python bench.py --config 2 --steps 2 --warmup 1 --bins-per-step 25 ) > gpurun_out/r2_b3_c2.log 2>&1; tail -c 1200 gpurun_out/r2_b3_c2.log | head -c 1200

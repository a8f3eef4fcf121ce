#!/bin/bash
# session check: Viterbi / filter / search parity tests, then a short default bench (stage times of one isolated batch)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python -m pytest tests/test_vitp_gpu.py tests/test_filters_gpu.py tests/test_search_gpu.py -x -q ) > gpurun_out/s3_vit.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s3_vit.log
python bench.py --steps 3 --warmup 3 --no-plugin --no-cpu-baseline > gpurun_out/s3_b32v.log 2> gpurun_out/s3_b32v.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/s3_b32v.log") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"])
print(d["gcups"]["stage_ms_per_step"]["isolated_batch"])
PY

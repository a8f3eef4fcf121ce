#!/bin/bash
# session check: every search-level parity test, then a short default bench (stage times of one isolated batch)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python -m pytest tests/test_search_gpu.py tests/test_text_parity_gpu.py tests/test_fuzz_gpu.py tests/test_align_gpu.py tests/test_find_e2e_gpu.py tests/test_fullsize_gpu.py -x -q ) > gpurun_out/s3_env.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s3_env.log
python bench.py --steps 3 --warmup 3 --no-plugin --no-cpu-baseline > gpurun_out/s3_b32e.log 2> gpurun_out/s3_b32e.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/s3_b32e.log") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["value"], d["ms_per_step"])
print(d["gcups"]["stage_ms_per_step"]["isolated_batch"])
PY

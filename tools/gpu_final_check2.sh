#!/bin/bash
# last check of the round: every GPU test, then the default bench line (outputs under gpurun_out/r2g_*)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2g_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_tests.log
( time python bench.py ) > gpurun_out/r2g_bench.log 2> gpurun_out/r2g_bench.err; echo "bench default rc=$?"; tail -c 300 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d = json.loads([x for x in open('gpurun_out/r2g_bench.log') if x.startswith('{')][-1])
print('value %.0f e2e %.0f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
print('   iso', d['gcups']['stage_ms_per_step']['isolated_batch'])
print('   plugin', d['plugin']['value'], d['plugin']['ms_per_bin'])
print('   cpu', {k: v for k, v in d['cpu_baseline'].items() if k not in ('sample', 'note')})
PY

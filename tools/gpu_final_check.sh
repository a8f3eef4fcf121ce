#!/bin/bash
# final check of the round on one B200: every GPU test, smoke, the default bench line + the reference arm, and the ncu launch
# list of the default command (per-launch times are cold-cache and serialised: compare shares).  Outputs under gpurun_out/r2f_*.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2f_smoke.log
( time python bench.py ) > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err; echo "bench default rc=$?"; tail -c 300 gpurun_out/r2f_bench.err
( time python bench.py --impl reference ) > gpurun_out/r2f_bench_ref.log 2>&1; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_default.csv python bench.py --steps 2 --warmup 1 --no-plugin --no-cpu-baseline > gpurun_out/r2f_launches_default.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import json
for f in ('r2f_bench', 'r2f_bench_ref'):
    try:
        d = json.loads([x for x in open('gpurun_out/%s.log' % f) if x.startswith('{')][-1])
        print(f, 'value %.0f e2e %.0f ms/step %.1f bins/step %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config'].get('bins_per_step')))
        if 'gcups' in d: print('   iso', d['gcups']['stage_ms_per_step']['isolated_batch'])
        if 'plugin' in d: print('   plugin', d['plugin']['value'], d['plugin']['ms_per_bin'])
        if 'cpu_baseline' in d: print('   cpu', {k: v for k, v in d['cpu_baseline'].items() if k not in ('sample', 'note')})
        if 'roofline' in d: print('   roofline', {k: d['roofline'][k] for k in ('achieved', 'peak', 'frac', 'kernel_ms_per_step')})
    except Exception as e:
        print(f, 'no line', e)
PY

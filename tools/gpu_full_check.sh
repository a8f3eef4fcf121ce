#!/bin/bash
# round-end check on one B200: GPU tests, smoke, default bench + reference arm, launch list of the default command, configs 4 / 2, diverse db, bin-statistics scan
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_t9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t9.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
( time python bench.py ) > gpurun_out/r2_b9.log 2> gpurun_out/r2_b9.err; echo "bench default rc=$?"; tail -c 300 gpurun_out/r2_b9.err
( time python bench.py --impl reference ) > gpurun_out/r2_b9_ref.log 2>&1; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_default.csv python bench.py --steps 2 --warmup 1 --no-plugin --no-cpu-baseline > gpurun_out/r2_launches_default.log 2>&1; echo "launch list rc=$?"
( time python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/r2_b9_c4.log 2>&1
( time python bench.py --db diverse --steps 3 --warmup 2 --no-plugin --no-cpu-baseline ) > gpurun_out/r2_b9_div.log 2>&1
( time python bench.py --config 2 --steps 2 --warmup 1 --bins-per-step 25 ) > gpurun_out/r2_b9_c2.log 2>&1
( time python tools/bench_binstats.py --bins 1000 ) > gpurun_out/r2_binstats_1000.json 2> gpurun_out/r2_binstats.err; echo "binstats rc=$?"; cut -c1-200 gpurun_out/r2_binstats_1000.json
python - <<'PY'
import json
for f in ('r2_b9','r2_b9_ref','r2_b9_c4','r2_b9_div','r2_b9_c2'):
    try:
        d=json.loads([x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')][-1])
        print(f,'value %.0f e2e %.0f ms/step %.1f bins/step %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config'].get('bins_per_step')))
        if 'gcups' in d: print('   iso',d['gcups']['stage_ms_per_step']['isolated_batch'])
        if 'plugin' in d: print('   plugin',d['plugin']['value'],d['plugin']['ms_per_bin'])
        if 'cpu_baseline' in d: print('   cpu',{k:v for k,v in d['cpu_baseline'].items() if k!='sample' and k!='note'})
        if 'cascade' in d: print('   cascade',d['cascade'])
    except Exception as e: print(f,'no line',e)
PY

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_t4.log
( time python bench.py --steps 3 --warmup 2 ) > gpurun_out/r2_b4.log 2> gpurun_out/r2_b4.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_b4.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_b4.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
    print('iso',d['gcups']['stage_ms_per_step']['isolated_batch'])
    print('plugin',d.get('plugin'))
    print('cpu',d.get('cpu_baseline'))
PY
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --pipeline 1 --bins-per-step 4 --no-plugin --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1
echo "launch list rc=$? lines $(wc -l < gpurun_out/r2_launches.csv)"
i=0
for k in "ssv_kernel<32" "vitp_kernel<2" "envelope2_kernel<8" "regions2_kernel<8" "fwd2_kernel<8" ; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$k" -c 1 -o gpurun_out/r2_ncu_$i -f python bench.py --steps 1 --warmup 0 --pipeline 1 --bins-per-step 4 --no-plugin --no-cpu-baseline > gpurun_out/r2_ncu_$i.log 2>&1
  echo "ncu $k rc=$?"
  ncu -i gpurun_out/r2_ncu_$i.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${i}_raw.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep

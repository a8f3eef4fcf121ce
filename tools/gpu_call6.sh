#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( python -m pytest tests/test_find_e2e_gpu.py tests/test_msv_gpu.py -m gpu -q ) > gpurun_out/r2_t6.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_t6.log
( python bench.py --steps 3 --warmup 2 ) > gpurun_out/r2_b6.log 2> gpurun_out/r2_b6.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_b6.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_b6.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
    print('plugin',d.get('plugin'))
PY
i=0
for k in 'ssv_kernel' 'vitp_kernel<\(int\)2,' 'envelope2_kernel<\(int\)8,' 'ensemble_kernel' 'fwd2_kernel<\(int\)8,' ; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$k" -c 1 -o gpurun_out/r2_ncu_$i -f python bench.py --steps 1 --warmup 0 --pipeline 1 --bins-per-step 4 --no-plugin --no-cpu-baseline > gpurun_out/r2_ncu_$i.log 2>&1
  echo "ncu $k rc=$?"
  ncu -i gpurun_out/r2_ncu_$i.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${i}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_ncu_$i.ncu-rep --page details --csv > gpurun_out/r2_ncu_${i}_details.csv 2>/dev/null
  [ $i -gt 1 ] && rm -f gpurun_out/r2_ncu_$i.ncu-rep
done
ls -la gpurun_out | head -30
( time python bench.py --config 4 --steps 1 --warmup 0 --no-cpu-baseline ) > gpurun_out/r2_b6_c4.log 2>&1; tail -c 900 gpurun_out/r2_b6_c4.log | head -c 900
du -sh gpurun_out

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python -m pytest tests/test_align_gpu.py tests/test_find_e2e_gpu.py tests/test_search_gpu.py -m gpu -q -s ) > gpurun_out/r2_t5.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_t5.log
( python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile-plugin ) > gpurun_out/r2_b5.log 2> gpurun_out/r2_b5.err; echo "bench rc=$?"
grep -A40 "cumulative" gpurun_out/r2_b5.err | cut -c1-150 | head -45
i=0
for k in 'ssv_kernel' 'vitp_kernel<\(int\)2,' 'envelope2_kernel<\(int\)8,' 'ensemble_kernel' 'fwd2_kernel<\(int\)8,' ; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$k" -c 1 -o gpurun_out/r2_ncu_$i -f python bench.py --steps 1 --warmup 0 --pipeline 1 --bins-per-step 4 --no-plugin --no-cpu-baseline > gpurun_out/r2_ncu_$i.log 2>&1
  echo "ncu $k rc=$? $(grep -c 'PROF.*Profiling' gpurun_out/r2_ncu_$i.log)"
  ncu -i gpurun_out/r2_ncu_$i.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${i}_raw.csv 2>/dev/null
  ls -la gpurun_out/r2_ncu_$i.ncu-rep
done

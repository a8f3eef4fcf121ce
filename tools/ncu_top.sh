#!/bin/bash
# ncu --set full captures of the top kernels of one bench step (development tool; run under gpurun).  Kernel names are matched
# on the demangled form, e.g. `void ckm::vitp_kernel<(int)2, (bool)0>(ckm::FilterParams)`.  Only the raw/details pages are kept
# (gpurun merges at most 64 MiB back); the first capture's .ncu-rep stays for `ncu -i ... --page source`.
mkdir -p gpurun_out
i=0
for k in 'ssv_kernel' 'vitp_kernel<\(int\)2,' 'envelope2_kernel<\(int\)8,' 'ensemble_kernel' 'fwd2_kernel<\(int\)8,' ; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$k" -c 1 -o gpurun_out/r2_ncu_$i -f \
      python bench.py --steps 1 --warmup 0 --pipeline 1 --bins-per-step 4 --no-plugin --no-cpu-baseline > gpurun_out/r2_ncu_$i.log 2>&1
  echo "ncu $k rc=$?"
  ncu -i gpurun_out/r2_ncu_$i.ncu-rep --page raw --csv > gpurun_out/r2_ncu_${i}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_ncu_$i.ncu-rep --page details --csv > gpurun_out/r2_ncu_${i}_details.csv 2>/dev/null
  [ $i -gt 1 ] && rm -f gpurun_out/r2_ncu_$i.ncu-rep
done

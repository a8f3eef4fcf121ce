#!/bin/bash
# ncu --set full captures of the top kernels of one bench step (development tool; run under gpurun)
mkdir -p gpurun_out
i=0
for k in "ssv_kernel<32" "msv_exact_kernel" "vit2_kernel<4" "envelope2_kernel<28" ; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$k" -c 1 -o gpurun_out/ncu_r1e_$i -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_r1e_$i.log 2>&1
  tail -2 gpurun_out/ncu_r1e_$i.log
done

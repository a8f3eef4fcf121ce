#!/bin/bash
# round-2 call 1: GPU tests + 8-warp crash hunt
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_t1.log
W8=$PWD/checkm_b200/libckm_w8.so
( CKM_LIBRARY=$W8 timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pipeline 2 ) > gpurun_out/r2_w8_plain.log 2>&1; echo "w8 plain rc=$?"
tail -3 gpurun_out/r2_w8_plain.log
( CKM_LIBRARY=$W8 timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pipeline 2 --bins-per-step 2 ) > gpurun_out/r2_w8_plain_b2.log 2>&1; echo "w8 plain b2 rc=$?"
tail -3 gpurun_out/r2_w8_plain_b2.log
( CKM_LIBRARY=$W8 timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --pipeline 2 --bins-per-step 2 ) > gpurun_out/r2_w8_memcheck.log 2>&1; echo "w8 memcheck rc=$?"
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r2_w8_memcheck.log; tail -4 gpurun_out/r2_w8_memcheck.log

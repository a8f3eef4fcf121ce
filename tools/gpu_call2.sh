#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|differ as text|rows whose text|all cases" gpurun_out/r2_t2.log | tail -20
W8=$PWD/checkm_b200/libckm_w8.so
for tool in racecheck initcheck; do
  for lib in "$W8" ""; do
    tag=$([ -n "$lib" ] && echo w8 || echo w16)
    ( CKM_LIBRARY=$lib timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_${tag}_${tool}.log 2>&1
    echo "$tag $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|smoke ok' gpurun_out/r2_${tag}_${tool}.log | tr '\n' ' ')"
  done
done

cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CKM_TRACE=1 python bench.py --steps 1 --warmup 1 --pipeline 1 --no-plugin --no-cpu-baseline > gpurun_out/r2_trace.log 2> gpurun_out/r2_trace.err
grep "ckm trace" gpurun_out/r2_trace.err | tail -24
for p in 1 3; do python bench.py --steps 3 --warmup 2 --pipeline $p --no-plugin --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('pipeline $p', round(d['value']), round(d['ms_per_step'],1))"; done

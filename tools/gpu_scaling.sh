#!/bin/bash
# 8 GPUs: weak scaling (config 3) and strong scaling of the fixed 512-bin set (config 4)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
( time $TR bench.py --gpus $N --steps 3 --warmup 2 ) > gpurun_out/r2_b8_n$N.log 2> gpurun_out/r2_b8_n$N.err; echo "n$N rc=$?"; tail -c 300 gpurun_out/r2_b8_n$N.err
( time $TR bench.py --gpus $N --config 4 --steps 2 --warmup 1 ) > gpurun_out/r2_b8_n${N}_c4.log 2> gpurun_out/r2_b8_n${N}_c4.err; echo "n$N c4 rc=$?"; tail -c 300 gpurun_out/r2_b8_n${N}_c4.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in ('r2_b8_n%s'%N,'r2_b8_n%s_c4'%N):
    try:
        d=json.loads([x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')][-1])
        print(f,'value %.0f e2e %.0f ms/step %.1f n_gpus %s bins/step %s per-gpu %s load %.1fs gen %.1fs'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['n_gpus'],d['config'].get('bins_per_step'),d['config'].get('per_gpu_bins_per_step'),d['config']['model_load_s'],d['config']['workload_generation_s']))
    except Exception as e: print(f,'no line',e)
PY

#!/bin/bash
# 2 GPUs: the multi-rank paths (NCCL QA gather through ckm_allgather_qa, LPT partition, reference arm on rank 0 only)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( python -m pytest tests/test_find_e2e_gpu.py tests/test_reduction_gpu.py -m gpu -q ) > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_t7.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
( time $TR bench.py --gpus 2 --steps 3 --warmup 2 ) > gpurun_out/r2_b7_n2.log 2> gpurun_out/r2_b7_n2.err; echo "n2 rc=$?"; tail -c 400 gpurun_out/r2_b7_n2.err
( time $TR bench.py --gpus 2 --config 4 --steps 1 --warmup 0 ) > gpurun_out/r2_b7_n2_c4.log 2> gpurun_out/r2_b7_n2_c4.err; echo "n2 c4 rc=$?"; tail -c 400 gpurun_out/r2_b7_n2_c4.err
( time $TR bench.py --gpus 2 --impl reference --steps 1 --warmup 0 ) > gpurun_out/r2_b7_n2_ref.log 2>&1; echo "n2 ref rc=$?"
( time python bench.py --steps 2 --warmup 1 --no-cpu-baseline ) > gpurun_out/r2_b7_n1.log 2>&1
python - <<'PY'
import json
for f in ('r2_b7_n2','r2_b7_n2_c4','r2_b7_n2_ref','r2_b7_n1'):
    try:
        d=json.loads([x for x in open('gpurun_out/%s.log'%f) if x.startswith('{')][-1])
        print(f,'value %.0f e2e %.0f ms/step %.1f n_gpus %s bins/step %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['n_gpus'],d['config'].get('bins_per_step')), d.get('plugin',{}).get('ms_per_bin'))
    except Exception as e: print(f,'no line',e)
PY

#!/bin/bash
# 2 GPUs: the cb_init tests (sharded DctHashIndex / CvFeaturesIndex vs oracle) and the bench under torchrun at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_similar_scale_gpu.py -q -m gpu -k "cb_init" > gpurun_out/r17_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r17_t.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 8 > gpurun_out/bench_t2.json 2> gpurun_out/bench_t2.err
echo "bench2 rc=$?" >> gpurun_out/bench_t2.err
CB_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 2 --warmup 3 --no-extras > /dev/null 2> gpurun_out/trace_t2.err
tail -n 3 gpurun_out/r17_t.log | cut -c1-300
tail -c 200 gpurun_out/bench_t2.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_t2.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('N=2 value %.3e step %.2f ms (kernel %.2f sort %.2f) e2e %.2f ms parity %s 100M %.1f ms %s'%(d['value'],d['ms_per_step'],r['kernel_ms_per_step'],r['sort_ms_per_step'],d['e2e']['ms_per_step'],d['parity'],d['target_100M']['ms_per_pass'],d['target_100M'].get('total',{}).get('consistent')))
except Exception as e: print('N=2 unreadable', e)
PY
grep "cb trace] rank 0" gpurun_out/trace_t2.err | tail -n 8

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r12_gpus.txt; nproc >> gpurun_out/r12_gpus.txt
timeout 600 python -m pytest tests/test_similar_scale_gpu.py -q -m gpu -k "cb_init" > gpurun_out/r12_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r12_t.log
for N in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 8 > gpurun_out/bench_s$N.json 2> gpurun_out/bench_s$N.err
  echo "bench$N rc=$?" >> gpurun_out/bench_s$N.err
done
timeout 600 python bench.py --gpus 1 --steps 8 --legs target_100M > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err
echo "bench1 rc=$?" >> gpurun_out/bench_s1.err
CB_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29550 bench.py --gpus 8 --steps 2 --warmup 3 --no-extras > /dev/null 2> gpurun_out/trace_n8.err
tail -n 3 gpurun_out/r12_t.log | cut -c1-300
for N in 1 2 4 8; do tail -c 120 gpurun_out/bench_s$N.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_s$N.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('N=$N value %.3e step %.2f ms (kernel %.2f sort %.2f) e2e %.2f ms parity %s 100M %.1f ms %s'%(d['value'],d['ms_per_step'],r['kernel_ms_per_step'],r['sort_ms_per_step'],d['e2e']['ms_per_step'],d['parity']['ok'],d['target_100M']['ms_per_pass'],d['target_100M'].get('total',{}).get('consistent')))
except Exception as e: print('N=$N unreadable', e)
PY
done
grep "cb trace] rank 0" gpurun_out/trace_n8.err | tail -n 8

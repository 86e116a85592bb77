#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mih_gpu.py tests/test_similar_scale_gpu.py -q -m gpu -x > gpurun_out/r9_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r9_t.log
timeout 600 python tools/mih_bench.py 1048576 3000000 10000000 --thr 5 --json gpurun_out/mih_bench_r02c.jsonl > gpurun_out/mih_bench_c.log 2>&1
CB_MIH2_SORT=1 timeout 300 python tools/mih_bench.py 10000000 --thr 5 > gpurun_out/mih_bench_sort.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err
echo "bench2 rc=$?" >> gpurun_out/bench_n2b.err
CB_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 2 --warmup 3 --no-extras > /dev/null 2> gpurun_out/trace_n2.err
timeout 600 python bench.py --steps 5 --legs target_100M > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err
tail -n 4 gpurun_out/r9_t.log | cut -c1-300
grep '"need": 2' gpurun_out/mih_bench_c.log | cut -c1-420; grep '"need": 2' gpurun_out/mih_bench_sort.log | cut -c1-420
for f in bench_n2b bench_n1b; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('$f', 'value %.3e step %.2f ms (kernel %.2f sort %.2f) e2e %.2f ms parity %s 100M %.1f ms'%(d['value'],d['ms_per_step'],r['kernel_ms_per_step'],r['sort_ms_per_step'],d['e2e']['ms_per_step'],d['parity']['ok'],d['target_100M']['ms_per_pass']))
except Exception as e: print('$f', 'unreadable', e)
PY
done
grep "cb trace" gpurun_out/trace_n2.err | tail -n 16

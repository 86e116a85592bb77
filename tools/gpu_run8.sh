#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r8_gpus.txt; nproc >> gpurun_out/r8_gpus.txt
timeout 600 python -m pytest tests/test_similar_scale_gpu.py -q -m gpu -k "cb_init" > gpurun_out/r8_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r8_t.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench8 rc=$?" >> gpurun_out/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/verify_100m.py --needles 0 --compare profiles/verify_100m_r02.json > gpurun_out/verify_100m_n8.json 2> gpurun_out/verify_100m_n8.err
echo "verify8 rc=$?" >> gpurun_out/verify_100m_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench4 rc=$?" >> gpurun_out/bench_n4.err
timeout 600 python bench.py --gpus 1 --legs target_100M > gpurun_out/bench_n1_8box.json 2> gpurun_out/bench_n1_8box.err
echo "bench1 rc=$?" >> gpurun_out/bench_n1_8box.err
tail -n 4 gpurun_out/r8_t.log | cut -c1-300
for f in bench_n8 bench_n4 bench_n1_8box; do tail -c 300 gpurun_out/$f.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', 'value %.3e step %.2f ms e2e %.3e %.2f ms parity %s 100M %.1f ms %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['parity']['ok'],d['target_100M']['ms_per_pass'],d['target_100M'].get('total')))
except Exception as e: print('$f', 'unreadable', e)
PY
done
cat gpurun_out/verify_100m_n8.json | cut -c1-900; tail -c 300 gpurun_out/verify_100m_n8.err

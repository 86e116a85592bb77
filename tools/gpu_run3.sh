#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
nvidia-smi -L > gpurun_out/r3_gpus.txt
timeout 600 python -m pytest tests/test_similar_scale_gpu.py -q -m gpu -k "cb_init or concurrent" > gpurun_out/r3_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3_t.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-extras > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench2 rc=$?" >> gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/verify_100m.py --rows 20000000 --needles 0 > gpurun_out/verify_20m_n2.json 2> gpurun_out/verify_20m_n2.err
echo "verify2 rc=$?" >> gpurun_out/verify_20m_n2.err
timeout 600 python tools/verify_100m.py --rows 20000000 --needles 4000 --out gpurun_out/verify_20m_n1.json > gpurun_out/verify_20m_n1.log 2>&1
echo "verify1 rc=$?" >> gpurun_out/verify_20m_n1.log
tail -5 gpurun_out/r3_t.log; tail -c 1500 gpurun_out/bench_n2.err; head -c 1200 gpurun_out/bench_n2.json; echo; cat gpurun_out/verify_20m_n2.json | cut -c1-600; tail -c 600 gpurun_out/verify_20m_n2.err; tail -2 gpurun_out/verify_20m_n1.log | cut -c1-900

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r6_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r6_t.log
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?" >> gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?" >> gpurun_out/bench_ref.err
tail -n 6 gpurun_out/r6_t.log | cut -c1-300; tail -c 1500 gpurun_out/bench_full.err; tail -c 300 gpurun_out/bench_ref.err; head -c 600 gpurun_out/bench_ref.json

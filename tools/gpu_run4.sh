#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mih_gpu.py tests/test_similar_scale_gpu.py -q -m gpu -x > gpurun_out/r4_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r4_t.log
timeout 600 python tools/mih_bench.py 1048576 3000000 10000000 --thr 5 --json gpurun_out/mih_bench_r02b.jsonl > gpurun_out/mih_bench_b.log 2>&1
echo "mih_bench rc=$?" >> gpurun_out/mih_bench_b.log
CB_MIH_WALK=1 timeout 300 python tools/mih_bench.py 10000000 --thr 5 > gpurun_out/mih_bench_walk.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --legs target_100M > gpurun_out/bench_q4.json 2> gpurun_out/bench_q4.err
echo "bench rc=$?" >> gpurun_out/bench_q4.err
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mih_gpu.py -x -q -m gpu -k "prefilter" > gpurun_out/r4_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r4_sanitizer.log
tail -8 gpurun_out/r4_t.log | cut -c1-300; grep '"need": 2' gpurun_out/mih_bench_b.log | cut -c1-500; tail -2 gpurun_out/mih_bench_walk.log | cut -c1-400; tail -c 400 gpurun_out/bench_q4.err; tail -3 gpurun_out/r4_sanitizer.log

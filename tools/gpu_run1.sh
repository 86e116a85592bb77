#!/bin/bash
# round-2 first GPU pass: parity of the new paths, variant timings, find queue, quick bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
nproc >> gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests/test_mih_gpu.py tests/test_dct_index_gpu.py tests/test_similar_scale_gpu.py -x -q -m gpu > gpurun_out/r2_t1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t1.log
timeout 600 python tools/mih_bench.py 1048576 10000000 --thr 5 --json gpurun_out/mih_bench_r02.jsonl > gpurun_out/mih_bench.log 2>&1
echo "mih_bench rc=$?" >> gpurun_out/mih_bench.log
timeout 120 ./cbird_b200/find_bench 1048576 32 2.0 5 > gpurun_out/find_bench.log 2>&1
echo "find_bench rc=$?" >> gpurun_out/find_bench.log
timeout 900 python bench.py --steps 5 --warmup 3 --legs target_100M > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
echo "bench rc=$?" >> gpurun_out/bench_q.err
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mih_gpu.py -x -q -m gpu -k "prefilter or segments or skewed" > gpurun_out/r2_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2_sanitizer.log
tail -5 gpurun_out/r2_t1.log; tail -3 gpurun_out/mih_bench.log; cat gpurun_out/find_bench.log; tail -c 600 gpurun_out/bench_q.err

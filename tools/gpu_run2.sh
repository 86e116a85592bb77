#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_mih_gpu.py tests/test_dct_index_gpu.py tests/test_similar_scale_gpu.py tests/test_scan_abi_gpu.py -q -m gpu > gpurun_out/r2_t2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_t2.log
timeout 600 python tools/mih_bench.py 1048576 3000000 10000000 --thr 5 --json gpurun_out/mih_bench_r02.jsonl > gpurun_out/mih_bench.log 2>&1
echo "mih_bench rc=$?" >> gpurun_out/mih_bench.log
for cfg in "2 25 16" "2 25 32" "3 25 32" "4 25 32" "2 5 32" "2 100 16" "1 25 32" "3 5 64"; do
  set -- $cfg
  echo "ctx=$1 spin_us=$2 threads=$3" >> gpurun_out/find_bench.log
  CB_FIND_CTX=$1 CB_FIND_SPIN_US=$2 timeout 120 ./cbird_b200/find_bench 1048576 $3 1.5 5 >> gpurun_out/find_bench.log 2>&1
done
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mih_gpu.py -x -q -m gpu -k "prefilter or segments or skewed" > gpurun_out/r2_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2_sanitizer.log
tail -8 gpurun_out/r2_t2.log; tail -4 gpurun_out/mih_bench.log | cut -c1-300; cat gpurun_out/find_bench.log | cut -c1-260; tail -3 gpurun_out/r2_sanitizer.log

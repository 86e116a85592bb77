#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mih_gpu.py tests/test_similar_scale_gpu.py tests/test_dct_index_gpu.py -q -m gpu -x > gpurun_out/r13_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r13_t.log
timeout 600 python tools/mih_bench.py 1048576 3000000 10000000 --thr 5 --json gpurun_out/mih_bench_r02e.jsonl > gpurun_out/mih_bench_e.log 2>&1
timeout 300 python tools/mih_bench.py 10000000 --thr 3,8 > gpurun_out/mih_bench_thr.log 2>&1
rm -f gpurun_out/find_bench3.log
for cfg in "3 5 32" "3 5 64" "4 2 64"; do
  set -- $cfg
  echo "ctx=$1 spin_us=$2 threads=$3" >> gpurun_out/find_bench3.log
  CB_FIND_CTX=$1 CB_FIND_SPIN_US=$2 timeout 120 ./cbird_b200/find_bench 1048576 $3 1.5 5 >> gpurun_out/find_bench3.log 2>&1
done
tail -n 4 gpurun_out/r13_t.log | cut -c1-300
grep '"need": 2' gpurun_out/mih_bench_e.log | cut -c1-440
cat gpurun_out/mih_bench_thr.log | cut -c1-300
cat gpurun_out/find_bench3.log | cut -c1-200

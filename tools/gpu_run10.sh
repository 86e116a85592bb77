#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r10_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r10_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r10_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r10_smoke.log
timeout 1500 python bench.py > gpurun_out/bench_full2.json 2> gpurun_out/bench_full2.err
echo "bench rc=$?" >> gpurun_out/bench_full2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mih2_bucket_kernel|mih2_scatter_kernel|mih2_hist_kernel|dct_hash32_kernel|find_small_kernel|similar_post_count" -c 14 -o gpurun_out/ncu_r02b -f python tools/profile_target_r02.py 10000000 > gpurun_out/ncu_r02b.log 2>&1
echo "ncu rc=$?" >> gpurun_out/ncu_r02b.log
rm -f gpurun_out/find_bench2.log
for cfg in "3 5 32" "4 5 32" "4 2 32" "3 0 32" "3 5 64" "4 2 64" "3 5 16"; do
  set -- $cfg
  echo "ctx=$1 spin_us=$2 threads=$3" >> gpurun_out/find_bench2.log
  CB_FIND_CTX=$1 CB_FIND_SPIN_US=$2 timeout 120 ./cbird_b200/find_bench 1048576 $3 1.5 5 >> gpurun_out/find_bench2.log 2>&1
done
timeout 2400 python tools/verify_100m.py --rows 100000000 --needles 100000 --out gpurun_out/verify_100m_r02.json > gpurun_out/verify_100m.log 2>&1
echo "verify rc=$?" >> gpurun_out/verify_100m.log
tail -n 5 gpurun_out/r10_t.log | cut -c1-300; tail -n 2 gpurun_out/r10_smoke.log; tail -c 400 gpurun_out/bench_full2.err; tail -n 2 gpurun_out/ncu_r02b.log; cat gpurun_out/find_bench2.log | cut -c1-200; tail -n 2 gpurun_out/verify_100m.log | cut -c1-900

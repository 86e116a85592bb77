#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r14_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r14_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r14_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r14_smoke.log
timeout 1500 python bench.py > gpurun_out/bench_full3.json 2> gpurun_out/bench_full3.err
echo "bench rc=$?" >> gpurun_out/bench_full3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mih2_bucket_kernel|mih2_scatter_all|mih2_hist_all|similar_post_count|similar_run_heads" -s 14 -c 9 -o gpurun_out/ncu_r02d -f python tools/profile_target_r02.py 10000000 > gpurun_out/ncu_r02d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dct_hash32_kernel|find_small_kernel|frame_hash_fused" -c 4 -o gpurun_out/ncu_r02e -f python tools/profile_target_r02.py 2000000 > gpurun_out/ncu_r02e.log 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mih_gpu.py -x -q -m gpu -k "prefilter or partition" > gpurun_out/r14_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r14_sanitizer.log
tail -n 4 gpurun_out/r14_t.log | cut -c1-300; tail -n 2 gpurun_out/r14_smoke.log; tail -c 300 gpurun_out/bench_full3.err; tail -n 1 gpurun_out/ncu_r02d.log; tail -n 1 gpurun_out/ncu_r02e.log; tail -n 3 gpurun_out/r14_sanitizer.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full3.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.3e step %.2f ms (kernel %.2f sort %.2f) e2e %.2f ms parity %s 100M %.1f ms'%(d['value'],d['ms_per_step'],r['kernel_ms_per_step'],r['sort_ms_per_step'],d['e2e']['ms_per_step'],d['parity']['ok'],d['target_100M']['ms_per_pass']))
print('find', d['find'].get('concurrent_find',{}).get('finds_per_s'), d['find'].get('concurrent_find_64',{}).get('finds_per_s'), 'flip', d['dct_hash'].get('flip_vs_cv2'))
PY

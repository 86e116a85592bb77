#!/bin/bash
# round-2 final single-GPU validation of the final build: GPU tests, smoke, the full bench line, launch list, one
# ncu --set full capture of a whole default-plan pass at 10^7 rows
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r18_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r18_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r18_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r18_smoke.log
timeout 1500 python bench.py > gpurun_out/bench_full5.json 2> gpurun_out/bench_full5.err
echo "bench rc=$?" >> gpurun_out/bench_full5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mih2_bucket_kernel|mih2_scatter_all_kernel|mih2_hist_all_kernel|similar_post_count|similar_run_heads|similar_post_scatter" -s 9 -c 9 -o gpurun_out/ncu_r02h -f python tools/profile_target_r02.py 10000000 > gpurun_out/ncu_r02h.log 2>&1
tail -n 4 gpurun_out/r18_t.log | cut -c1-300; tail -n 2 gpurun_out/r18_smoke.log; tail -c 300 gpurun_out/bench_full5.err; tail -n 1 gpurun_out/ncu_r02h.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full5.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.3e step %.2f ms (kernel %.2f sort %.2f) e2e %.2f ms parity %s 100M %.1f ms'%(d['value'],d['ms_per_step'],r['kernel_ms_per_step'],r['sort_ms_per_step'],d['e2e']['ms_per_step'],d['parity']['ok'],d['target_100M']['ms_per_pass']))
print('find', d['find'].get('concurrent_find',{}).get('finds_per_s'), d['find'].get('concurrent_find_64',{}).get('finds_per_s'), 'flip', d['dct_hash'].get('flip_vs_cv2'))
PY

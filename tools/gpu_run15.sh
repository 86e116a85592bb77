#!/bin/bash
# round-2 final single-GPU validation: GPU tests, smoke, the full bench line, launch list, ncu captures of the kernels
# the bench's roofline objects cite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r15_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r15_t.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r15_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r15_smoke.log
timeout 1500 python bench.py > gpurun_out/bench_full4.json 2> gpurun_out/bench_full4.err
echo "bench rc=$?" >> gpurun_out/bench_full4.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref4.json 2> gpurun_out/bench_ref4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mih2_scatter_all_kernel|mih2_hist_all_kernel" -s 2 -c 2 -o gpurun_out/ncu_r02f -f python tools/profile_target_r02.py 10000000 > gpurun_out/ncu_r02f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dct_hash32_kernel|frame_hash_fused_kernel" -s 2 -c 2 -o gpurun_out/ncu_r02g -f python tools/profile_target_r02.py 1000000 > gpurun_out/ncu_r02g.log 2>&1
tail -n 4 gpurun_out/r15_t.log | cut -c1-300; tail -n 2 gpurun_out/r15_smoke.log; tail -c 300 gpurun_out/bench_full4.err; tail -n 1 gpurun_out/ncu_r02f.log; tail -n 1 gpurun_out/ncu_r02g.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_full4.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.3e step %.2f ms (kernel %.2f sort %.2f) e2e %.2f ms parity %s 100M %.1f ms'%(d['value'],d['ms_per_step'],r['kernel_ms_per_step'],r['sort_ms_per_step'],d['e2e']['ms_per_step'],d['parity']['ok'],d['target_100M']['ms_per_pass']))
print('find', d['find'].get('concurrent_find',{}).get('finds_per_s'), d['find'].get('concurrent_find_64',{}).get('finds_per_s'), 'flip', d['dct_hash'].get('flip_vs_cv2'))
PY

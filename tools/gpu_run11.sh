#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mih_gpu.py tests/test_similar_scale_gpu.py tests/test_scan_abi_gpu.py -q -m gpu -x > gpurun_out/r11_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r11_t.log
timeout 600 python tools/mih_bench.py 1048576 3000000 10000000 --thr 5 --json gpurun_out/mih_bench_r02d.jsonl > gpurun_out/mih_bench_d.log 2>&1
timeout 600 python bench.py --steps 5 --legs target_100M > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dct_hash32_kernel|find_small_kernel|similar_post_count|mih2_scatter_kernel|frame_hash_fused" -c 10 -o gpurun_out/ncu_r02c -f python tools/profile_target_r02.py 10000000 > gpurun_out/ncu_r02c.log 2>&1
tail -n 4 gpurun_out/r11_t.log | cut -c1-300
grep '"need": 2' gpurun_out/mih_bench_d.log | cut -c1-420
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n1c.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.3e step %.2f ms (kernel %.2f sort %.2f) e2e %.2f ms parity %s 100M %.1f ms'%(d['value'],d['ms_per_step'],r['kernel_ms_per_step'],r['sort_ms_per_step'],d['e2e']['ms_per_step'],d['parity']['ok'],d['target_100M']['ms_per_pass']))
PY
tail -n 2 gpurun_out/ncu_r02c.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mih2_bucket_kernel -s 2 -c 3 -o gpurun_out/ncu_mih2_r02 -f python tools/profile_target_r02.py 10000000 > gpurun_out/ncu_mih2.log 2>&1
echo "ncu1 rc=$?" >> gpurun_out/ncu_mih2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mih_bucket_kernel -s 1 -c 1 -o gpurun_out/ncu_mihb_r02 -f python tools/profile_target_r02.py 10000000 > gpurun_out/ncu_mihb.log 2>&1
echo "ncu2 rc=$?" >> gpurun_out/ncu_mihb.log
tail -n 3 gpurun_out/ncu_mih2.log; tail -n 3 gpurun_out/ncu_mihb.log
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nproc > gpurun_out/r7_nproc.txt
timeout 2400 python tools/verify_100m.py --rows 100000000 --needles 20000 --out gpurun_out/verify_100m_r02.json > gpurun_out/verify_100m.log 2>&1
echo "verify rc=$?" >> gpurun_out/verify_100m.log
tail -n 3 gpurun_out/verify_100m.log | cut -c1-1200

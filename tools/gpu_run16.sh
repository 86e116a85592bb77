#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_mih_gpu.py tests/test_similar_scale_gpu.py tests/test_dct_index_gpu.py tests/test_cpp_adapter.py -x -q -m gpu > gpurun_out/r16_t.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r16_t.log
timeout 200 python tools/step_profile.py 10000000 > gpurun_out/r16_step.json 2> gpurun_out/r16_step.err
timeout 200 python tools/step_profile.py 3000000 >> gpurun_out/r16_step.json 2>> gpurun_out/r16_step.err
tail -n 3 gpurun_out/r16_t.log | cut -c1-300; cat gpurun_out/r16_step.json; tail -n 3 gpurun_out/r16_step.err

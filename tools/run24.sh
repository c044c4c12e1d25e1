#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/r24.log
t() { for i in 1 2 3 4; do env "$@" timeout 120 python -m pytest tests/test_gpu_boundary.py -q -k two_handles 2>&1 | grep -E "^E .*(diverged|differs)|passed|failed" | cut -c1-200 | tr '\n' ' ' | sed "s/^/[$*] /" >> gpurun_out/r24.log; echo >> gpurun_out/r24.log; done; }
t X=1
t B2S_TEST_MODE=2
t B2S_TEST_MODE=0
t B2S_TEST_SEQ=1
t B2S_GROUPS=1
t B2S_NO_GRAPH=1
t B2S_GRAPH_PER_GROUP=0
t B2S_CTRL_SPLIT=0
t B2S_TIER_SMALL=96,288
t CUDA_DEVICE_MAX_CONNECTIONS=32
cat gpurun_out/r24.log

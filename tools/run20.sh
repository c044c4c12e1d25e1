#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/r20_*
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_boundary.py -q 2>&1 | grep -E "^E .*(diverged|differs)|passed|failed" | cut -c1-300 >> gpurun_out/r20_file.log; done
for i in 1 2 3 4 5; do timeout 300 python -m pytest tests/test_gpu_boundary.py -q -k two_handles 2>&1 | grep -E "^E .*(diverged|differs)|passed|failed" | cut -c1-300 >> gpurun_out/r20_alone.log; done
echo file; cat gpurun_out/r20_file.log; echo alone; cat gpurun_out/r20_alone.log

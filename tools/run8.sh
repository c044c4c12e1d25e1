#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T="python -m pytest tests/test_gpu_task_logic.py -x -q -k free_run -s"
t() { name=$1; shift; env "$@" timeout 300 $T > gpurun_out/r8_$name.log 2>&1; echo "$name exit $?" >> gpurun_out/r8_summary.log; tail -3 gpurun_out/r8_$name.log | cut -c1-200 >> gpurun_out/r8_summary.log; }
rm -f gpurun_out/r8_summary.log
t default B2S_X=1
t lb256x2 B2S_LIB=robosuite_b200/variants/libb2s_lb256x2.so
t notier B2S_TIER_SMALL=96,288
t nosplit B2S_CTRL_SPLIT=0
t nostage B2S_NO_STAGE=1
t nograph B2S_NO_GRAPH=1
t G1 B2S_GROUPS=1
cat gpurun_out/r8_summary.log

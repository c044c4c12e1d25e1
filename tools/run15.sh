#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/r15_*.log
T="python -m pytest tests/test_gpu_boundary.py -q -k two_handles"
for i in 1 2 3 4; do timeout 300 $T >> gpurun_out/r15_default.log 2>&1; done
for i in 1 2 3 4; do B2S_LIB=robosuite_b200/variants/libb2s_zero.so timeout 300 $T >> gpurun_out/r15_zero.log 2>&1; done
for i in 1 2 3; do B2S_GROUPS=4 timeout 300 $T >> gpurun_out/r15_G4.log 2>&1; done
for i in 1 2 3; do B2S_GROUPS=1 timeout 300 $T >> gpurun_out/r15_G1.log 2>&1; done
for i in 1 2 3; do B2S_TIER_SMALL=96,288 timeout 300 $T >> gpurun_out/r15_notier.log 2>&1; done
for i in 1 2 3; do B2S_CTRL_SPLIT=0 timeout 300 $T >> gpurun_out/r15_nosplit.log 2>&1; done
for i in 1 2 3; do B2S_NO_GRAPH=1 timeout 300 $T >> gpurun_out/r15_nograph.log 2>&1; done
for i in 1 2 3; do B2S_NO_STAGE=1 timeout 300 $T >> gpurun_out/r15_nostage.log 2>&1; done
for f in gpurun_out/r15_*.log; do echo "$f: $(grep -cE '^1 passed' $f) passed, $(grep -cE '^1 failed' $f) failed"; done
timeout 900 python -m pytest tests/test_gpu_reset.py -x -q > gpurun_out/r15_reset.log 2>&1; tail -15 gpurun_out/r15_reset.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python tools/debug_two_handles.py 2>&1 | tail -12; echo ---; done > gpurun_out/r16_dbg.log 2>&1
for i in 1 2; do NSUB=1 timeout 300 python tools/debug_two_handles.py 2>&1 | tail -12; echo ---; done > gpurun_out/r16_dbg_nsub1.log 2>&1
cat gpurun_out/r16_dbg.log gpurun_out/r16_dbg_nsub1.log
timeout 900 python -m pytest tests/test_gpu_reset.py tests/test_gpu_env.py -x -q > gpurun_out/r16_reset.log 2>&1; tail -15 gpurun_out/r16_reset.log
timeout 600 python tools/probe_unstable.py PickPlace Panda 1024 250 f32 > gpurun_out/r16_unstable_pp.log 2>&1; tail -12 gpurun_out/r16_unstable_pp.log
timeout 600 python tools/probe_unstable.py NutAssemblyRound Panda 1024 250 f32 > gpurun_out/r16_unstable_nut.log 2>&1; tail -12 gpurun_out/r16_unstable_nut.log

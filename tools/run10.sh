#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B2S_GROUPS=4 B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r10_instr_Lift.log 2>&1
cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r10_instr_Lift.json
timeout 600 python -m pytest tests/test_gpu_task_logic.py -q -s -k lockstep > gpurun_out/r10_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r10_pytest.log
echo done

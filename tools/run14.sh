#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r14_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r14_pytest.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_boundary.py -q -k two_handles >> gpurun_out/r14_two_handles.log 2>&1; done
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-timeline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r14_b_$name.json 2> gpurun_out/r14_b_$name.err; }
run G8 B2S_GROUPS=8
run G16 B2S_GROUPS=16
B2S_GROUPS=8 B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r14_instr_Lift.log 2>&1
cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r14_instr_Lift_G8.json
echo done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B2S_GROUPS=8 B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r11_instr_Lift.log 2>&1
cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r11_instr_Lift.json
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-timeline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r11_b_$name.json 2> gpurun_out/r11_b_$name.err; }
run G4 B2S_GROUPS=4
run G8 B2S_GROUPS=8
run G16 B2S_GROUPS=16
run G32 B2S_GROUPS=32
run G8_cvx3552 B2S_GROUPS=8 B2S_CVX_BLOCKS=3552
run G8_cvx1024 B2S_GROUPS=8 B2S_CVX_BLOCKS=1024
run G16_onegraph B2S_GROUPS=16 B2S_GRAPH_PER_GROUP=0
timeout 600 python -m pytest tests/test_gpu_boundary.py -q -s > gpurun_out/r11_pytest_boundary.log 2>&1; echo "pytest exit $?" >> gpurun_out/r11_pytest_boundary.log
echo done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-timeline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r12_b_$name.json 2> gpurun_out/r12_b_$name.err; }
run G8_c8 B2S_GROUPS=8 CUDA_DEVICE_MAX_CONNECTIONS=8
run G8_c32 B2S_GROUPS=8
run G16_c32 B2S_GROUPS=16
run G32_c32 B2S_GROUPS=32
run G16_c32_onegraph B2S_GROUPS=16 B2S_GRAPH_PER_GROUP=0
run G4_c32 B2S_GROUPS=4
B2S_GROUPS=16 B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r12_instr_Lift.log 2>&1
cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r12_instr_Lift_G16.json
echo done

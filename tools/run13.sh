#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r13_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r13_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-timeline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r13_b_$name.json 2> gpurun_out/r13_b_$name.err; }
run G8 B2S_GROUPS=8
run G16 B2S_GROUPS=16
run G4 B2S_GROUPS=4
run G32 B2S_GROUPS=32
B2S_GROUPS=8 B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r13_instr_Lift.log 2>&1
cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r13_instr_Lift_G8.json
timeout 900 python bench.py --steps 10 --warmup 3 --config 3 --no-timeline --no-cpu-baseline > gpurun_out/r13_bench_c3.json 2> gpurun_out/r13_bench_c3.err
timeout 900 python bench.py --steps 10 --warmup 3 --config 5 --no-timeline --no-cpu-baseline > gpurun_out/r13_bench_c5.json 2> gpurun_out/r13_bench_c5.err
echo done

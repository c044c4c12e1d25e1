#!/bin/bash
# round 2, GPU run 3: per-kernel layouts + capacity tiers + descriptor slots + split controller: full suite, A/B benches, timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B2S_VERBOSE=1 timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-timeline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r3_b_$name.json 2> gpurun_out/r3_b_$name.err; }
run default B2S_X=1
run notier B2S_TIER_SMALL=32,64
run nosplit B2S_CTRL_SPLIT=0
run inline B2S_CTRL_FORK=0
for g in 2 8 16; do run G$g B2S_GROUPS=$g; done
for v in lb224x4 lb256x4 lb256x2; do run $v B2S_LIB=robosuite_b200/variants/libb2s_$v.so; run ${v}_G8 B2S_LIB=robosuite_b200/variants/libb2s_$v.so B2S_GROUPS=8; done
B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r3_instr_Lift.log 2>&1
cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r3_instr_Lift.json
B2S_GROUPS=8 B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r3_instr_Lift_G8.log 2>&1
cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r3_instr_Lift_G8.json
for spec in "Stack Sawyer 8192 JOINT_VELOCITY" "NutAssemblyRound Panda 4096 OSC_POSE" "Door Panda 2048 OSC_POSE" "PickPlace Panda 2048 OSC_POSE" "Stack Panda 2048 OSC_POSE"; do
  B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 600 python tools/probe_instr.py $spec > gpurun_out/r3_instr_$(echo $spec | cut -d' ' -f1-2 | tr ' ' _).log 2>&1
done
timeout 900 python bench.py --steps 10 --warmup 3 --config 3 --no-timeline --no-cpu-baseline > gpurun_out/r3_bench_c3.json 2> gpurun_out/r3_bench_c3.err
timeout 900 python bench.py --steps 6 --warmup 3 --config 4 --no-timeline --no-cpu-baseline > gpurun_out/r3_bench_c4.json 2> gpurun_out/r3_bench_c4.err
timeout 900 python bench.py --steps 10 --warmup 3 --config 5 --no-timeline --no-cpu-baseline > gpurun_out/r3_bench_c5.json 2> gpurun_out/r3_bench_c5.err
timeout 300 python tools/probe_reset.py Lift 4096 > gpurun_out/r3_probe_reset.log 2>&1
echo done

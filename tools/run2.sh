#!/bin/bash
# round 2, GPU run 2: full GPU suite with the split controller kernel + A/B of controller placement and block sizes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-timeline"
B2S_CTRL_SPLIT=0 timeout 300 $B > gpurun_out/r2_b_nosplit.json 2>gpurun_out/r2_b_nosplit.err
timeout 300 $B > gpurun_out/r2_b_split_fork.json 2>gpurun_out/r2_b_split_fork.err
B2S_CTRL_FORK=0 timeout 300 $B > gpurun_out/r2_b_split_inline.json 2>gpurun_out/r2_b_split_inline.err
for w in 8 10 12 13; do B2S_WARPS_PER_BLOCK=$w timeout 300 $B > gpurun_out/r2_b_split_wpb$w.json 2>gpurun_out/r2_b_split_wpb$w.err; done
for g in 2 3 6 8; do B2S_GROUPS=$g timeout 300 $B > gpurun_out/r2_b_split_G$g.json 2>gpurun_out/r2_b_split_G$g.err; done
B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r2_instr_Lift.log 2>&1
cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r2_instr_Lift_split.json
B2S_WARPS_PER_BLOCK=12 B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r2_instr_Lift12.log 2>&1
cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r2_instr_Lift_split_wpb12.json
timeout 600 python bench.py --steps 10 --warmup 3 --config 5 --no-timeline --no-cpu-baseline > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
timeout 300 python tools/probe_reset.py Lift 4096 > gpurun_out/r2_probe_reset.log 2>&1
echo done

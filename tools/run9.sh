#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r9_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-timeline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r9_b_$name.json 2> gpurun_out/r9_b_$name.err; }
run default B2S_X=1
run G8 B2S_GROUPS=8
for cfgv in "G4 B2S_GROUPS=4" "G8 B2S_GROUPS=8"; do
  set -- $cfgv; nm=$1; shift
  env "$@" B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r9_instr_Lift_$nm.log 2>&1
  cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r9_instr_Lift_$nm.json
done
B2S_GROUPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase1_kernel -s 30 -c 3 -o gpurun_out/r9_prof_p1 python tools/probe_pipeline.py Lift Panda 4096 OSC_POSE 2 > gpurun_out/r9_ncu_p1.log 2>&1
B2S_GROUPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tail_kernel -s 60 -c 2 -o gpurun_out/r9_prof_tail python tools/probe_pipeline.py Lift Panda 4096 OSC_POSE 2 > gpurun_out/r9_ncu_tail.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --config 3 --no-timeline --no-cpu-baseline > gpurun_out/r9_bench_c3.json 2> gpurun_out/r9_bench_c3.err
timeout 900 python bench.py --steps 10 --warmup 3 --config 5 --no-timeline --no-cpu-baseline > gpurun_out/r9_bench_c5.json 2> gpurun_out/r9_bench_c5.err
echo done

#!/bin/bash
# round 2, GPU run 7: EPA horizon in shared memory, per-group graphs, bigger small tier; Stack crash under compute-sanitizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r7_memcheck.log python -m pytest tests/test_gpu_task_logic.py -x -q -k "free_run and Stack" > gpurun_out/r7_memcheck_pytest.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r7_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r7_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-timeline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r7_b_$name.json 2> gpurun_out/r7_b_$name.err; }
run default B2S_X=1
run onegraph B2S_GRAPH_PER_GROUP=0
for g in 2 8 16 32; do run G$g B2S_GROUPS=$g; done
run G8_onegraph B2S_GROUPS=8 B2S_GRAPH_PER_GROUP=0
run lb256x2_G8 B2S_LIB=robosuite_b200/variants/libb2s_lb256x2.so B2S_GROUPS=8
run lb256x2_G16 B2S_LIB=robosuite_b200/variants/libb2s_lb256x2.so B2S_GROUPS=16
run nosplit_G8 B2S_CTRL_SPLIT=0 B2S_GROUPS=8
for cfgv in "G4 B2S_GROUPS=4" "G8 B2S_GROUPS=8" "G16 B2S_GROUPS=16"; do
  set -- $cfgv; nm=$1; shift
  env "$@" B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r7_instr_Lift_$nm.log 2>&1
  cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r7_instr_Lift_$nm.json
done
timeout 900 python bench.py --steps 10 --warmup 3 --config 3 --no-timeline --no-cpu-baseline > gpurun_out/r7_bench_c3.json 2> gpurun_out/r7_bench_c3.err
timeout 900 python bench.py --steps 10 --warmup 3 --config 5 --no-timeline --no-cpu-baseline > gpurun_out/r7_bench_c5.json 2> gpurun_out/r7_bench_c5.err
echo done

#!/bin/bash
# round 2, GPU run 6: warp-uniform tier decision + single phase-1 kernel (no graph forks): full suite, A/B benches, timelines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r6_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-timeline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r6_b_$name.json 2> gpurun_out/r6_b_$name.err; }
run default B2S_X=1
run notier B2S_TIER_SMALL=96,288
run nosplit B2S_CTRL_SPLIT=0
run nostage B2S_NO_STAGE=1
for g in 2 8 16; do run G$g B2S_GROUPS=$g; done
for v in lb224x4 lb256x4 lb256x2; do run $v B2S_LIB=robosuite_b200/variants/libb2s_$v.so; run ${v}_G8 B2S_LIB=robosuite_b200/variants/libb2s_$v.so B2S_GROUPS=8; done
for cfgv in "A B2S_X=1" "G8 B2S_GROUPS=8"; do
  set -- $cfgv; nm=$1; shift
  env "$@" B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r6_instr_Lift_$nm.log 2>&1
  cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r6_instr_Lift_$nm.json
done
timeout 900 python bench.py --steps 10 --warmup 3 --config 3 --no-timeline --no-cpu-baseline > gpurun_out/r6_bench_c3.json 2> gpurun_out/r6_bench_c3.err
timeout 900 python bench.py --steps 6 --warmup 3 --config 4 --no-timeline --no-cpu-baseline > gpurun_out/r6_bench_c4.json 2> gpurun_out/r6_bench_c4.err
timeout 900 python bench.py --steps 10 --warmup 3 --config 5 --no-timeline --no-cpu-baseline > gpurun_out/r6_bench_c5.json 2> gpurun_out/r6_bench_c5.err
timeout 300 python tools/probe_reset.py Lift 4096 > gpurun_out/r6_probe_reset.log 2>&1
echo done

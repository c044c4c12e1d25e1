#!/bin/bash
# round 2, GPU run 4: localise the small-tier discrepancy; concurrency / launch-bound A/B without tiers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/debug_tier.py > gpurun_out/r4_debug_tier.log 2>&1
export B2S_TIER_SMALL=96,288
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-timeline"
run() { name=$1; shift; env "$@" timeout 300 $B > gpurun_out/r4_b_$name.json 2> gpurun_out/r4_b_$name.err; }
run notier B2S_X=1
run notier_nostage B2S_NO_STAGE=1
run notier_nocfork B2S_CTRL_FORK=0
run notier_nonfork B2S_NARROW_FORK=0
run notier_noforks B2S_CTRL_FORK=0 B2S_NARROW_FORK=0
run notier_nosplit B2S_CTRL_SPLIT=0 B2S_NARROW_FORK=0
run notier_G8 B2S_GROUPS=8
run notier_G2 B2S_GROUPS=2
for v in lb224x4 lb256x4 lb256x2; do run notier_$v B2S_LIB=robosuite_b200/variants/libb2s_$v.so; done
for cfgv in "A B2S_X=1" "B B2S_CTRL_FORK=0 B2S_NARROW_FORK=0" "C B2S_CTRL_SPLIT=0 B2S_NARROW_FORK=0"; do
  set -- $cfgv; nm=$1; shift
  env "$@" B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 300 python tools/probe_instr.py Lift Panda 4096 OSC_POSE > gpurun_out/r4_instr_Lift_$nm.log 2>&1
  cp gpurun_out/instr_Lift_Panda_4096.json gpurun_out/r4_instr_Lift_$nm.json
done
timeout 300 python tools/probe_reset.py Lift 4096 > gpurun_out/r4_probe_reset.log 2>&1
echo done

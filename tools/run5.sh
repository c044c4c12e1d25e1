#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for t in 8,32 8,64 32,32 16,48 4,24 7,32 8,31 8,33; do timeout 200 python tools/debug_tier.py $t > gpurun_out/r5_tier_$t.log 2>&1; done
B2S_GROUPS=1 timeout 200 python tools/debug_tier.py 8,32 > gpurun_out/r5_tier_8,32_G1.log 2>&1
B2S_WPB5=1 timeout 200 python tools/debug_tier.py 8,32 > gpurun_out/r5_tier_8,32_wpb1.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r5_memcheck.log python tools/debug_tier.py 8,32 > gpurun_out/r5_tier_memcheck.log 2>&1
tail -3 gpurun_out/r5_tier_*.log
echo done

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/r23.log
t() { for i in 1 2 3; do env "${@:3}" timeout 120 python tools/debug_two_handles2.py $1 $2 2>&1 | tail -1 | sed "s/^/[${*:3}] /" >> gpurun_out/r23.log; done; }
t 1 conc X=1
t 2 conc X=1
t 0 conc X=1
t 1 seq X=1
t 1 swap X=1
t 1 onlyB X=1
t 1 conc B2S_GROUPS=1
t 1 conc B2S_NO_GRAPH=1
t 1 conc CUDA_DEVICE_MAX_CONNECTIONS=32
t 1 conc B2S_CTRL_SPLIT=0
t 1 conc B2S_TIER_SMALL=96,288
cat gpurun_out/r23.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for n in 64 512 4096; do echo "== n $n barriers 1"; B2S_UNIT_BARRIERS=1 timeout 40 python tools/probe_unit.py $n 3 2>&1 | tail -4 | cut -c1-250; done 2>&1 | tee gpurun_out/r28_probe.log

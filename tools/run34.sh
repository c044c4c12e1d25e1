#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
for k in tail_kernel phase1_kernel phase0_kernel; do
  timeout 170 ncu --set full --clock-control none -k regex:$k -s 410 -c 1 -f -o /tmp/r34_$k python tools/probe_pipeline.py Lift Panda 4096 OSC_POSE 3 > gpurun_out/r34_$k.log 2>&1
  ncu -i /tmp/r34_$k.ncu-rep --page raw --csv > gpurun_out/r34_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/r34_$k.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/r34_${k}_sass.csv.gz
  ls -la /tmp/r34_$k.ncu-rep gpurun_out/r34_${k}_raw.csv gpurun_out/r34_${k}_sass.csv.gz
done
du -sh gpurun_out

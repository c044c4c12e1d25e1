#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.0f e2e %.0f ms %.2f warn %s launches %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["config"]["solver_warn_flags"],d["gpu_launches"]))'
run() { nm=$1; shift; timeout 150 python bench.py "$@" > gpurun_out/r39_$nm.json 2> gpurun_out/r39_$nm.err; tail -1 gpurun_out/r39_$nm.json | python -c "$P" $nm 2>&1 | tail -1; }
run c2_full --steps 20 --warmup 3
run c5 --config 5 --steps 6 --warmup 3 --no-cpu-baseline --no-timeline --preroll 40

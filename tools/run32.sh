#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/r32_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r32_pytest.log; tail -6 gpurun_out/r32_pytest.log | cut -c1-200
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.0f e2e %.0f ms %.2f warn %s launches %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["config"]["solver_warn_flags"],d["gpu_launches"]))'
run() { nm=$1; shift; timeout 170 python bench.py "$@" > gpurun_out/r32_$nm.json 2> gpurun_out/r32_$nm.err; tail -1 gpurun_out/r32_$nm.json | python -c "$P" $nm 2>&1 | tail -1; }
run c2_m1_full --steps 20 --warmup 3
run c2_m2 --steps 20 --warmup 3 --mode 2 --no-cpu-baseline --no-timeline
run c3_m1 --config 3 --steps 8 --warmup 3 --no-cpu-baseline --no-timeline --preroll 60
run c3_m2 --config 3 --steps 8 --warmup 3 --no-cpu-baseline --no-timeline --preroll 60 --mode 2
run c5_m1 --config 5 --steps 8 --warmup 3 --no-cpu-baseline --no-timeline --preroll 60
run c5_m2 --config 5 --steps 8 --warmup 3 --no-cpu-baseline --no-timeline --preroll 60 --mode 2
run c4_m1 --config 4 --steps 4 --warmup 3 --no-cpu-baseline --no-timeline --preroll 20
run c4_m2 --config 4 --steps 4 --warmup 3 --no-cpu-baseline --no-timeline --preroll 20 --mode 2

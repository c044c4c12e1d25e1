#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r36_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r36_pytest.log; tail -4 gpurun_out/r36_pytest.log | cut -c1-300; grep -E "^FAILED|^E  " gpurun_out/r36_pytest.log | head -10 | cut -c1-250
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.0f e2e %.0f ms %.2f warn %s launches %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["config"]["solver_warn_flags"],d["gpu_launches"]))'
run() { nm=$1; shift; timeout 150 python bench.py "$@" > gpurun_out/r36_$nm.json 2> gpurun_out/r36_$nm.err; tail -1 gpurun_out/r36_$nm.json | python -c "$P" $nm 2>&1 | tail -1; }
run c2 --steps 10 --warmup 3 --no-cpu-baseline --no-timeline
run c5 --config 5 --steps 6 --warmup 3 --no-cpu-baseline --no-timeline --preroll 40

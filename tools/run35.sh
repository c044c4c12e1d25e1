#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/r35_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r35_pytest.log; tail -4 gpurun_out/r35_pytest.log | cut -c1-200
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.0f e2e %.0f ms %.2f warn %s launches %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["config"]["solver_warn_flags"],d["gpu_launches"]))'
run() { nm=$1; shift; timeout 200 python bench.py "$@" > gpurun_out/r35_$nm.json 2> gpurun_out/r35_$nm.err; tail -1 gpurun_out/r35_$nm.json | python -c "$P" $nm 2>&1 | tail -1; }
run c2_m1_full --steps 20 --warmup 3
run c2_m2 --steps 20 --warmup 3 --mode 2 --no-cpu-baseline --no-timeline
run c3_m1 --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-timeline
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

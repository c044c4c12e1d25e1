#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -k "unit_queue" -x > gpurun_out/r22_unit_tests.log 2>&1; tail -15 gpurun_out/r22_unit_tests.log | cut -c1-250
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-timeline --preroll 60"
for md in 1 2; do
  B2S_VERBOSE=1 timeout 300 $B --mode $md 2> gpurun_out/r22_m$md.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mode $md value %.0f e2e %.0f ms %.2f warn %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['solver_warn_flags']))"
  grep "unit-queue" gpurun_out/r22_m$md.err | head -2
done 2>&1 | tee gpurun_out/r22_modes.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_boundary.py -q -k two_handles 2>&1 | grep -E "^E .*(diverged|differs)|passed|failed" | cut -c1-300; done | tee gpurun_out/r22_two.log

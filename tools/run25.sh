#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for i in 1 2 3 4; do timeout 120 python -m pytest tests/test_gpu_boundary.py -q -k two_handles 2>&1 | grep -E "^E .*(diverged|differs)|passed|failed" | cut -c1-200; done | tee gpurun_out/r25_two.log
timeout 600 python -m pytest tests/test_gpu_engine.py -q -k "unit_queue" -x > gpurun_out/r25_unit_tests.log 2>&1; tail -5 gpurun_out/r25_unit_tests.log | cut -c1-250
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-timeline --preroll 60 --mode 2"
for bar in 4 3 2 1 0; do
  B2S_UNIT_BARRIERS=$bar timeout 300 $B 2> gpurun_out/r25_b$bar.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('barriers $bar value %.0f e2e %.0f ms %.2f warn %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['solver_warn_flags']))"
done 2>&1 | tee gpurun_out/r25_modes.log
for w in 4 6; do
  B2S_UNIT_WPB=$w timeout 300 $B 2> gpurun_out/r25_w$w.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wpb $w value %.0f e2e %.0f ms %.2f warn %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['solver_warn_flags']))"
done 2>&1 | tee -a gpurun_out/r25_modes.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for bar in 0 4; do echo "== barriers $bar"; B2S_UNIT_BARRIERS=$bar timeout 60 python tools/probe_unit.py 16 2 2>&1 | tail -4 | cut -c1-250; echo "exit $?"; done 2>&1 | tee gpurun_out/r26_probe.log
if grep -q "^ok" gpurun_out/r26_probe.log; then
  timeout 150 python -m pytest tests/test_gpu_engine.py -q -k "unit_queue" -x > gpurun_out/r26_unit_tests.log 2>&1; tail -5 gpurun_out/r26_unit_tests.log | cut -c1-250
  B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-timeline --preroll 40 --mode 2"
  for bar in 4 2 0; do
    B2S_UNIT_BARRIERS=$bar timeout 100 $B 2> gpurun_out/r26_b$bar.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('barriers $bar value %.0f e2e %.0f ms %.2f warn %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['solver_warn_flags']))" 2>&1 | tail -1
  done 2>&1 | tee gpurun_out/r26_modes.log
fi

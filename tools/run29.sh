#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== n 64 barriers 4"; B2S_UNIT_BARRIERS=4 timeout 40 python tools/probe_unit.py 64 3 2>&1 | tail -2 | cut -c1-250
for bar in 4 0; do echo "== tests barriers $bar"; B2S_UNIT_BARRIERS=$bar timeout 100 python -m pytest tests/test_gpu_engine.py -q -k "unit_queue" 2>&1 | tail -3 | cut -c1-200; done 2>&1 | tee gpurun_out/r29_tests.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-timeline --preroll 30 --mode 2"
for bar in 4 2 1 0; do
  B2S_UNIT_BARRIERS=$bar timeout 70 $B 2> gpurun_out/r29_b$bar.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('barriers $bar value %.0f e2e %.0f ms %.2f warn %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['solver_warn_flags']))" 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r29_modes.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-timeline --preroll 30 --mode 2"
run() { nm=$1; shift; env "$@" B2S_VERBOSE=1 timeout 70 $B 2> gpurun_out/r30_$nm.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$nm value %.0f e2e %.0f ms %.2f warn %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['config']['solver_warn_flags']))" 2>&1 | tail -1; grep "unit-queue" gpurun_out/r30_$nm.err | head -1; }
run u512_b4 B2S_LIB=robosuite_b200/variants/libb2s_u512.so B2S_UNIT_BARRIERS=4
run u512_b4_w12 B2S_LIB=robosuite_b200/variants/libb2s_u512.so B2S_UNIT_BARRIERS=4 B2S_UNIT_WPB=12
run b4_148blocks B2S_UNIT_BARRIERS=4 B2S_UNIT_BLOCKS=148
run b4_w4 B2S_UNIT_BARRIERS=4 B2S_UNIT_WPB=4

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-timeline --preroll 50"
run() { nm=$1; shift; env "$@" timeout 60 $B 2> gpurun_out/r37_$nm.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$nm value %.0f ms %.2f warn %s'%(d['value'],d['ms_per_step'],d['config']['solver_warn_flags']))" 2>&1 | tail -1; }
run G8 B2S_GROUPS=8
run G6 B2S_GROUPS=6
run G12 B2S_GROUPS=12
run G16 B2S_GROUPS=16
run G8_cvx128 B2S_GROUPS=8 B2S_CVX_BLOCKS=128
run G8_cvx512 B2S_GROUPS=8 B2S_CVX_BLOCKS=512
run G8_wpb5_4 B2S_GROUPS=8 B2S_WPB5=4
run G8_wpb0_4 B2S_GROUPS=8 B2S_WPB0=4

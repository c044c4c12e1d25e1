#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-timeline --preroll 60"
for sc in 1 2 4; do for G in 8 16; do
  B2S_BENCH_SCALE=$sc B2S_GROUPS=$G timeout 300 $B 2> gpurun_out/r21_s${sc}_G$G.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('scale $sc G $G value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))"
done; done 2>&1 | tee gpurun_out/r21_scale.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/r20_*
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_boundary.py -q 2>&1 | grep -E "^E .*(diverged|differs)|passed|failed" | cut -c1-300 >> gpurun_out/r20_file.log; done
for i in 1 2 3 4 5; do timeout 300 python -m pytest tests/test_gpu_boundary.py -q -k two_handles 2>&1 | grep -E "^E .*(diverged|differs)|passed|failed" | cut -c1-300 >> gpurun_out/r20_alone.log; done
echo file; cat gpurun_out/r20_file.log; echo alone; cat gpurun_out/r20_alone.log
#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-timeline --preroll 60"
for sc in 1 2 4; do for G in 8 16; do
  B2S_BENCH_SCALE=$sc B2S_GROUPS=$G timeout 300 $B 2> gpurun_out/r21_s${sc}_G$G.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('scale $sc G $G value %.0f e2e %.0f ms %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step']))"
done; done 2>&1 | tee gpurun_out/r21_scale.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/r33_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r33_pytest.log; tail -4 gpurun_out/r33_pytest.log | cut -c1-200
for t in PickPlace Door Stack NutAssemblyRound; do
  B2S_GROUPS=8 B2S_LIB=robosuite_b200/variants/libb2s_instr.so timeout 120 python tools/probe_instr.py $t Panda 2048 OSC_POSE > gpurun_out/r33_instr_$t.log 2>&1
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/instr_${t}_Panda_2048.json'))
    s=d['solver']
    print('$t','step_ms %.1f'%d['step_ms_events'],'kernels',{k:round(v['mean_us']) for k,v in d['kernels'].items()},'large_tier',s['large_tier_env_substeps'],'of',s['solves'],'ncon',s['ncon'],'nefc',s['nefc'],'niter %.2f'%s['mean_niter'],'warn',d['warn'])
except Exception as e: print('$t failed',e)
PY
done
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.0f e2e %.0f ms %.2f warn %s launches %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["config"]["solver_warn_flags"],d["gpu_launches"]))'
run() { nm=$1; shift; timeout 170 python bench.py "$@" > gpurun_out/r33_$nm.json 2> gpurun_out/r33_$nm.err; tail -1 gpurun_out/r33_$nm.json | python -c "$P" $nm 2>&1 | tail -1; }
run c5_m1 --config 5 --steps 8 --warmup 3 --no-cpu-baseline --no-timeline --preroll 60

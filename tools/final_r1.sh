timeout 600 python -m pytest tests -q -m gpu -s 2>&1 | grep -E "^E  |passed|failed|FAILED|vs reference" | cut -c1-260 | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1g.json; cut -c1-330 gpurun_out/bench_r1g.json; echo
for v in bar2 bar4; do
  echo "$v: $(B2S_LIB=robosuite_b200/libb2s_$v.so timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"value": [0-9.]*' | head -2 | tr '\n' ' ')"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:phase_kernel -s 241 -c 1 -o gpurun_out/prof_r1g python tools/probe_pipeline.py > gpurun_out/r1g.log 2>&1; ls -la gpurun_out/prof_r1g.ncu-rep

bash tools/sweep_lmem.sh
timeout 300 python -m pytest tests/test_gpu_env.py -q -m gpu -k gym 2>&1 | tail -3
for lib in robosuite_b200/libb2s.so build_variants/libb2s_small.so; do
  echo "$lib: $(B2S_LIB=$lib timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"value": [0-9.]*' | head -2 | tr '\n' ' ')"
done

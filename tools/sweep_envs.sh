for cfg in "1024 1" "2048 2" "2048 1" "4096 4" "4096 2" "8192 4" "8192 8"; do set -- $cfg
  echo "envs=$1 groups=$2: $(B2S_BENCH_ENVS=$1 B2S_GROUPS=$2 timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' | head -2 | tr '\n' ' ')"
done

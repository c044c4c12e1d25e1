timeout 500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for cfg in "14 4" "7 4" "7 8" "2 4" "2 8" "14 8"; do set -- $cfg
  echo "wpb=$1 groups=$2: $(B2S_WARPS_PER_BLOCK=$1 B2S_GROUPS=$2 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"value": [0-9.]*' | head -2 | tr '\n' ' ')"
done

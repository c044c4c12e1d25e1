for v in "" _vb _vc ""; do
  lib=robosuite_b200/libb2s$v.so
  echo "$lib: $(B2S_LIB=$lib timeout 200 python bench.py --steps 12 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"value": [0-9.]*' | head -2 | tr '\n' ' ')"
done
for v in _vb _vc; do
  echo "tests $v: $(B2S_LIB=robosuite_b200/libb2s$v.so timeout 400 python -m pytest tests -q -m gpu 2>&1 | tail -1)"
done

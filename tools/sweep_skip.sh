for cfg in "1 0" "1 1" "1 2" "1 3" "4 0" "4 3"; do set -- $cfg
  echo "groups=$1 skip=$2: $(B2S_GROUPS=$1 B2S_DEBUG_SKIP=$2 timeout 200 python tools/probe_graphstep.py 4096 2>&1 | tail -1)"
done

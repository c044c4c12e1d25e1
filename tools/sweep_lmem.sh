for cfg in "1 1024 0" "1 1024 1" "4 4096 0" "4 4096 1" "8 4096 0"; do set -- $cfg
  if [ "$3" = "1" ]; then export B2S_NO_LMEM_FLAG=1; else unset B2S_NO_LMEM_FLAG; fi
  echo "groups=$1 envs=$2 noflag=$3: $(B2S_VERBOSE=1 B2S_GROUPS=$1 timeout 200 python tools/probe_graphstep.py $2 2>&1 | tail -2 | tr '\n' ' ')"
done

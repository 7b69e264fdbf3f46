#!/bin/bash
# A/B harness with a value check: tools/ab2.sh "<variant>[:ENV=VAL,...]" ...   (run under gpurun)
# prints the c3 logp+grad time and the first logp values (must agree between variants)
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  res=$(env $envs PTM_B200_LIB=$PWD/ab/lib_$v.so timeout 60 python tools/quick_bench.py 1000000 1024 ${AB_DTYPE:-f64} 2>&1 | grep -E "logp\+grad|tensor|Error|error" | sed 's/evals.*//' | tr '\n' ' ')
  echo "$spec -> $res"
done

#!/bin/bash
# A/B harness: tools/ab.sh "<variant>[:ENV=VAL,...]" ...   (run under gpurun)
for spec in "$@"; do
  v=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  res=$(env $envs PTM_B200_LIB=$PWD/ab/lib_$v.so timeout 100 python tools/quick_bench.py 1000000 1024 2>&1 | grep "logp+grad" | sed 's/evals.*//')
  echo "$spec -> $res"
done

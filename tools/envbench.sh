#!/bin/bash
# tools/envbench.sh "ENV=VAL,ENV=VAL" ...  -> c3 logp+grad time of the in-tree library under each setting
for spec in "$@"; do
  envs=$(echo "$spec" | tr ',' ' ')
  res=$(env $envs timeout 90 python tools/quick_bench.py 1000000 1024 ${AB_DTYPE:-f64} 2>&1 | grep -E "logp\+grad|Error|error" | sed 's/evals.*//' | tr '\n' ' ')
  echo "$spec -> $res"
done

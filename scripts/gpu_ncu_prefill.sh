#!/bin/bash
# ncu --set full captures of the prompt-batch kernels: usage bash scripts/gpu_ncu_prefill.sh TAG config "kernel-regex:skip:count" ...
TAG=$1; CFG=$2; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
export B200_TMP=/tmp/b200_models
for spec in "$@"; do
  IFS=: read -r rx skip cnt <<< "$spec"
  name=$(echo $rx | tr -c 'a-zA-Z0-9_\n' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -f -o $OUT/prof_${name}_$skip \
      python scripts/ncu_prefill.py $CFG > $OUT/ncu_${name}.log 2>&1
  tail -2 $OUT/ncu_${name}.log
done
ls -la $OUT

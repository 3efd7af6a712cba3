#!/bin/bash
# ncu --set full capture of ONE launch of the persistent per-token kernel (8B Q4_K_M, n_kv ~ 2000)
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export B200_TMP=/tmp/b200_models
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_token' -s 3 -c 1 \
    -f -o $OUT/prof_token python scripts/ncu_token.py 5 > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT

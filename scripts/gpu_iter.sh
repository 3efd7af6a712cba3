#!/bin/bash
# A/B visit: quick parity subset, then value-only bench lines under different env settings, then a phase trace.
# usage: bash scripts/gpu_iter.sh tag "ENV1=.. ENV2=.." "ENV.." ...   (each further argument = one bench configuration)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
export B200_TMP=/tmp/b200_models
if [ -z "$SKIP_TESTS" ]; then
  ( time timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_decode_parity_gpu.py -m gpu -x -q \
      -k "quantize or mul_mat_vec_golden or rms_norm or golden_models or graph_equals or long_context or fullshape_twins" ) > $OUT/pytest_subset.log 2>&1
  tail -4 $OUT/pytest_subset.log
fi
i=0
for cfg in "$@"; do
  i=$((i+1))
  echo "== cfg$i: $cfg" | tee -a $OUT/ab.txt
  ( env $cfg timeout 600 python bench.py --value-only --steps ${STEPS:-6} 2>> $OUT/bench.err ) | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: continue
    print(round(j['value'], 1), 'tok/s', j['clocks'])
" | tee -a $OUT/ab.txt
done
if [ -n "$TRACE_CFG" ]; then
  env $TRACE_CFG timeout 300 python scripts/trace_token.py 2000 $OUT/trace_token.txt > $OUT/trace_head.txt 2>&1; head -9 $OUT/trace_head.txt
fi

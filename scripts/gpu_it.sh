#!/bin/bash
# one development visit: a pytest selection, value-only bench lines (one per further argument = env settings), a phase trace
# usage: bash scripts/gpu_it.sh TAG "pytest -k expression or empty" "ENV=.. ENV=.." ...
TAG=$1; KEXPR=$2; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
export B200_TMP=/tmp/b200_models
if [ -n "$KEXPR" ]; then
  ( time timeout ${TEST_TIMEOUT:-900} python -m pytest tests -m gpu -x -q --timeout ${PER_TEST_TIMEOUT:-180} --timeout-method=thread -k "$KEXPR" ) > $OUT/pytest.log 2>&1
  tail -12 $OUT/pytest.log
fi
i=0
for cfg in "$@"; do
  i=$((i+1))
  echo "== cfg$i: $cfg" | tee -a $OUT/ab.txt
  ( env $cfg timeout 600 python bench.py --value-only --steps ${STEPS:-6} $BENCH_ARGS 2>> $OUT/bench.err ) | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: continue
    print(round(j['value'], 1), 'tok/s', j['clocks'], j['config']['name'])
" | tee -a $OUT/ab.txt
done
tail -5 $OUT/bench.err 2>/dev/null
if [ -n "$TRACE" ]; then
  timeout 300 python scripts/trace_token.py 2000 $OUT/trace_token.txt > $OUT/trace_head.txt 2>&1; head -24 $OUT/trace_head.txt
fi

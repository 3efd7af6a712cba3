#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one ncu --set full capture of a layer's kernels.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export B200_TMP=/tmp/b200_models
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.csv 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
if [ -z "$SKIP_TESTS" ]; then
  ( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
  tail -5 $OUT/pytest_gpu.log
fi
if [ -z "$SKIP_BENCH" ]; then
  ( time timeout 900 python bench.py $BENCH_ARGS ) > $OUT/bench.json 2> $OUT/bench.err
  cat $OUT/bench.json
  ( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_reference.json 2>> $OUT/bench.err
  cut -c1-400 $OUT/bench_reference.json
fi
if [ -z "$SKIP_TRACE" ]; then
  timeout 300 python scripts/trace_token.py 2000 $OUT/trace_token.txt > $OUT/trace_head.txt 2>&1; head -12 $OUT/trace_head.txt
fi
if [ -n "$AB" ]; then
  BOOSTER_B200_NO_PDL=1 timeout 900 python bench.py --no-cpu > $OUT/bench_nopdl.json 2>> $OUT/bench.err; cat $OUT/bench_nopdl.json
  BOOSTER_B200_ATTN_SPLIT=1 timeout 900 python bench.py --no-cpu > $OUT/bench_attnsplit.json 2>> $OUT/bench.err; cat $OUT/bench_attnsplit.json
fi
if [ -z "$SKIP_NCU" ]; then
  # every launch of 3 un-graphed tokens with its device time (cold-cache, serialised: shares, not absolutes)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_matvec|k_attn|k_embed|k_argmax' -c 600 \
      --csv --log-file $OUT/launches.csv python scripts/ncu_token.py 3 > $OUT/ncu_launches.log 2>&1
  # layer 0 of token 2: QKV, scores, softmax+P.V, wo, gate/up, down
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_matvec|k_attn' -s 193 -c 6 \
      -f -o $OUT/prof_layer python scripts/ncu_token.py 2 > $OUT/ncu_full.log 2>&1
  tail -3 $OUT/ncu_full.log
  python scripts/launch_summary.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1 || true
fi
ls -la $OUT

#!/bin/bash
# round-2 visit A: all GPU parity tests, the default bench line, two more BASELINE configs
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export B200_TMP=/tmp/b200_models
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/smi.csv 2>&1
nproc > $OUT/nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -15 $OUT/pytest_gpu.log
( time timeout 900 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
cat $OUT/bench.json; tail -3 $OUT/bench.err
( time timeout 900 python bench.py --config 8b-q8_0-prefill512 --steps 2 --warmup 1 --no-cpu ) > $OUT/bench_q8_0.json 2>> $OUT/bench.err
cat $OUT/bench_q8_0.json; tail -3 $OUT/bench.err
( time timeout 900 python bench.py --config mistral-q5km-8192 --steps 4 --warmup 3 --no-cpu ) > $OUT/bench_mistral.json 2>> $OUT/bench.err
cat $OUT/bench_mistral.json; tail -3 $OUT/bench.err
ls -la $OUT

OUT=gpurun_out/r01h; mkdir -p $OUT
export B200_TMP=/tmp/b200_models
python scripts/trace_token.py 2000 $OUT/trace_token.txt > $OUT/trace_head.txt 2>&1
BOOSTER_B200_DUP=1 python scripts/trace_token.py 2000 $OUT/trace_dup.txt > $OUT/trace_dup_head.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log
python bench.py --no-cpu > $OUT/bench.json 2>$OUT/bench.err; cut -c1-200 $OUT/bench.json

#!/bin/bash
# prefill visit: speed of the prompt batch on configs, ncu launch list of the first layers of one prompt batch
# usage: bash scripts/gpu_prefill.sh TAG "config config ..." [pytest -k expr]
TAG=$1; CFGS=$2; KEXPR=$3
OUT=gpurun_out/$TAG
mkdir -p $OUT
export B200_TMP=/tmp/b200_models
if [ -n "$KEXPR" ]; then
  ( time timeout ${TEST_TIMEOUT:-900} python -m pytest tests -m gpu -x -q --timeout 300 --timeout-method=thread -k "$KEXPR" ) > $OUT/pytest.log 2>&1
  tail -12 $OUT/pytest.log
fi
for cfg in $CFGS; do
  timeout 600 python scripts/prefill_speed.py $cfg 512 3 2>> $OUT/err.txt | tee -a $OUT/speed.txt
  if [ -n "$AB_ENV" ]; then env $AB_ENV timeout 600 python scripts/prefill_speed.py $cfg 512 3 2>> $OUT/err.txt | sed "s/^/[$AB_ENV] /" | tee -a $OUT/speed.txt; fi
  if [ -z "$SKIP_NCU" ]; then
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_quant_batch|k_matmul_batch|k_mma_batch|k_umma_batch|k_attn|k_embed_batch' -c ${NCU_C:-80} --csv --log-file $OUT/launches_$cfg.csv \
        python scripts/ncu_prefill.py $cfg > $OUT/ncu_$cfg.log 2>&1
    python scripts/launch_summary.py $OUT/launches_$cfg.csv | tee $OUT/launches_${cfg}_summary.txt
  fi
done
tail -5 $OUT/err.txt

#!/bin/bash
# 2-GPU visit: layer-split bench with the NVLink peer hand-off and with NCCL send/recv (value-only lines)
TAG=${1:-n2}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export B200_TMP=/tmp/b200_models
for mode in 1 0; do
  echo "== P2P=$mode N=$N" | tee -a $OUT/n.txt
  BOOSTER_B200_P2P=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 6 --warmup 3 --value-only $BENCH_ARGS 2>> $OUT/err.txt | tee $OUT/bench_p2p${mode}.json | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: continue
    print(round(j['value'], 1), 'tok/s', j['config']['hand_off'][:40], 'crc', j['burst_ids_crc32'], 'e2e', round(j['e2e']['value'],1))
" | tee -a $OUT/n.txt
done
tail -5 $OUT/err.txt

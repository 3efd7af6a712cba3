#!/bin/bash
# bench lines of the BASELINE GPU configs: bash scripts/gpu_configs.sh TAG "config ..."
TAG=$1; CFGS=$2
OUT=gpurun_out/$TAG
mkdir -p $OUT
export B200_TMP=/tmp/b200_models
for cfg in $CFGS; do
  extra="--no-cpu"; steps="--steps 4 --warmup 3"
  [ "$cfg" = "8b-q4km-2048" ] && { extra=""; steps=""; }
  [ "$cfg" = "8b-q8_0-prefill512" ] && steps="--steps 2 --warmup 3"
  ( time timeout 900 python bench.py --config $cfg $steps $extra ) > $OUT/bench_$cfg.json 2>> $OUT/bench.err
  python - <<EOF
import json
for l in open("$OUT/bench_$cfg.json"):
    try: j = json.loads(l)
    except Exception: continue
    print("$cfg", "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "token_roofline", round(j["token_roofline"]["frac_of_peak"], 3),
          "roofline.frac", round(j["roofline"]["frac"], 3), "prompt_batch", j.get("prompt_batch", {}).get("tokens_per_s"), "prefill", j.get("prefill", {}).get("tokens_per_s"),
          "doInference", j["e2e"].get("doInference", {}))
EOF
done
tail -5 $OUT/bench.err

#!/bin/bash
# One gpurun call: lane kernel geometry sweep (warps per CTA) on the full headline batch and on C3 / C5.
TAG=${1:-wsweep}
WS=${2:-"14 16 20 24"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for w in $WS; do
  BROTLI_B200_LANE_WARPS=$w timeout 600 python bench.py --steps 3 --warmup 3 --unique 2048 --no-e2e --no-cpu --no-other-configs > $OUT/bench_w$w.json 2> $OUT/bench_w$w.err
  python -c "import json; j=json.load(open('$OUT/bench_w$w.json')); print('W=$w headline', j['value'], 'GB/s ms', j['ms_per_step'], 'bit_exact', j.get('bit_exact'))"
  BROTLI_B200_LANE_WARPS=$w timeout 900 python profiles/gpu_configs.py > $OUT/configs_w$w.jsonl 2> $OUT/configs_w$w.err
  python -c "
import json
for l in open('$OUT/configs_w$w.jsonl'):
    j=json.loads(l); print('W=$w', j['config'], j['GBps'], 'GB/s', j['ms'], 'ms bailed', j['bailed_to_exact'], j['bit_exact'])"
done

#!/bin/bash
# One gpurun call: table-slot size (E) against warps per SM, full headline batch.  E = (slot bytes - 16) / 2.
TAG=${1:-eslot}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for cfg in "14 372" "14 308" "14 220" "16 308" "16 220" "20 220" "20 180"; do
  set -- $cfg
  BROTLI_B200_LANE_WARPS=$1 BROTLI_B200_LANE_SLOT_BYTES=$2 timeout 600 python bench.py --steps 3 --warmup 2 --unique 2048 --no-e2e --no-cpu --no-other-configs > $OUT/bench_w$1_s$2.json 2> $OUT/bench_w$1_s$2.err
  python -c "import json; j=json.load(open('$OUT/bench_w$1_s$2.json')); print('W=$1 slot=$2 headline', j['value'], 'GB/s ms', j['ms_per_step'], 'bit_exact', j.get('bit_exact'))"
done

#!/bin/bash
# One gpurun call: lane kernel time against batch size in units of one wave (resident lanes) for several geometries:
# separates the per-wave cost from the partial last wave.   gpurun -- 'bash profiles/gpu_waves.sh r05w "14 16 20 24"'
TAG=${1:-waves}
WS=${2:-"14 16 20 24"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for w in $WS; do
  lanes=$((148 * w * 32))
  for mult in 100 200 300 77 177; do
    n=$((lanes * mult / 100))
    BROTLI_B200_LANE_WARPS=$w timeout 600 python bench.py --streams $n --steps 3 --warmup 2 --unique 2048 --no-e2e --no-cpu --no-other-configs > $OUT/bench_w${w}_m$mult.json 2> $OUT/bench_w${w}_m$mult.err
    python -c "import json; j=json.load(open('$OUT/bench_w${w}_m$mult.json')); print('W=$w waves=$mult% n=$n', j['value'], 'GB/s ms', j['ms_per_step'], 'streams/ms', round($n/j['ms_per_step'],1), 'bit_exact', j.get('bit_exact'))"
  done
done

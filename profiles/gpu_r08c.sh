#!/bin/bash
# One gpurun call: measurement-only ablations of the lane kernel's memory instructions (BD_LANE_PROBE_ABLATE), headline probe.
TAG=${1:-r08c}
VARS=${2:-"default ab1 ab2 ab3 ab7"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for rep in 1 2; do
for v in $VARS; do
  LIB=""
  if [ "$v" != "default" ]; then LIB=$PWD/rust-brotli-decompressor_b200/variants/libbrotli_b200_$v.so; fi
  BROTLI_B200_LIB=$LIB timeout 300 python bench.py --streams 189440 --unique 2048 --steps 3 --warmup 3 --no-e2e --no-cpu --no-other-configs > $OUT/bench_${v}_$rep.json 2> $OUT/bench_${v}_$rep.err
  python -c "import json; j=json.load(open('$OUT/bench_${v}_$rep.json')); print('$v rep $rep headline probe', j['value'], 'GB/s ms', j['ms_per_step'], 'bit_exact', j.get('bit_exact'))"
done
done 2>&1 | tee $OUT/summary.txt

#!/bin/bash
# One gpurun call: config C4 at its real shape (4096 x 16 MiB, lgwin 24: the warp-per-stream kernel) with library variants.
TAG=${1:-c4}
VARS=${2:-"default"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in $VARS; do
  LIB=""
  if [ "$v" != "default" ]; then LIB=$PWD/rust-brotli-decompressor_b200/variants/libbrotli_b200_$v.so; fi
  BROTLI_B200_LIB=$LIB timeout 900 python bench.py --config C4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-other-configs > $OUT/bench_c4_$v.json 2> $OUT/bench_c4_$v.err
  python -c "import json; j=json.load(open('$OUT/bench_c4_$v.json')); print('$v C4', j['value'], 'GB/s ms', j['ms_per_step'], 'bit_exact', j.get('bit_exact'))"
done

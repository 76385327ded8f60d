#!/bin/bash
# One gpurun call: source-level ncu capture (one launch of 94 720 streams) of the default build and one ablation; hot lines by stall samples.
TAG=${1:-r08i}
VARS=${2:-"default ab10"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in $VARS; do
  LIB=""
  if [ "$v" != "default" ]; then LIB=$PWD/rust-brotli-decompressor_b200/variants/libbrotli_b200_$v.so; fi
  BROTLI_B200_LIB=$LIB timeout 500 ncu --set full --clock-control none --import-source on -k regex:brotli_decode_lane -s 3 -c 1 -o $OUT/cap_$v -f \
    python bench.py --streams 94720 --unique 2048 --steps 1 --warmup 3 --no-e2e --no-cpu --no-other-configs > $OUT/cap_$v.log 2>&1
  python profiles/ncu_hot.py $OUT/cap_$v.ncu-rep > $OUT/hot_$v.txt 2>&1
  ncu -i $OUT/cap_$v.ncu-rep --page source --csv --print-source sass > $OUT/sass_$v.csv 2>/dev/null
  rm -f $OUT/cap_$v.ncu-rep
  head -20 $OUT/hot_$v.txt
done

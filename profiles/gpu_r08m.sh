#!/bin/bash
# One gpurun call: latency configuration of the small lane geometries (more literals per round) against the throughput
# configuration, batches below one wave (8 warps per SM) -- what an 8-GPU split of the headline batch hands each GPU.
TAG=${1:-r08m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in ${VARS:-default nolat}; do
  LIB=""
  if [ "$v" != "default" ]; then LIB=$PWD/rust-brotli-decompressor_b200/variants/libbrotli_b200_$v.so; fi
  LAT_TAG=$v BROTLI_B200_LIB=$LIB timeout 400 python profiles/gpu_latency.py 6144,8192,16384,32768,37888 > $OUT/lat_$v.jsonl 2> $OUT/lat_$v.err
  cat $OUT/lat_$v.jsonl
done 2>&1 | tee $OUT/summary.txt

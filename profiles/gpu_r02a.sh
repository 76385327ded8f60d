#!/bin/bash
# round 2, first GPU call: baseline A/B (default vs hpr) and the batch-size latency table under three geometries
OUT=gpurun_out/r02a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
for v in default hpr; do
  LIB=""; if [ "$v" != "default" ]; then LIB=$PWD/rust-brotli-decompressor_b200/variants/libbrotli_b200_$v.so; fi
  for rep in 1 2; do
    BROTLI_B200_LIB=$LIB timeout 600 python bench.py --streams 131072 --unique 2048 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/bench_${v}_$rep.json 2> $OUT/bench_${v}_$rep.err
    python -c "import json; j=json.load(open('$OUT/bench_${v}_$rep.json')); print('$v rep $rep', j['value'], 'GB/s ms', j['ms_per_step'], j.get('bit_exact'))"
  done
done
LAT_TAG=w14 timeout 600 python profiles/gpu_latency.py > $OUT/lat_w14.jsonl 2> $OUT/lat_w14.err; cat $OUT/lat_w14.jsonl
LAT_TAG=w8 BROTLI_B200_LANE_WARPS=8 timeout 600 python profiles/gpu_latency.py > $OUT/lat_w8.jsonl 2> $OUT/lat_w8.err; cat $OUT/lat_w8.jsonl
LAT_TAG=w4 BROTLI_B200_LANE_WARPS=4 timeout 600 python profiles/gpu_latency.py 1,32,256,1024,4096,8192,16384 > $OUT/lat_w4.jsonl 2> $OUT/lat_w4.err; cat $OUT/lat_w4.jsonl
LAT_TAG=exact BROTLI_B200_LANE=0 timeout 600 python profiles/gpu_latency.py 1,32,256,1024,4096,8192 > $OUT/lat_exact.jsonl 2> $OUT/lat_exact.err; cat $OUT/lat_exact.jsonl
tail -3 $OUT/*.err

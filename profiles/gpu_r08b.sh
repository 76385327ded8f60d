#!/bin/bash
# One gpurun call: DRAM traffic (ncu, one launch of 94 720 streams = one wave at 20 warps per SM) of the default build against
# the two measurement-only probes whose copy sources come from the lane's most recent output (near: 20..35 bytes back,
# near64: 64..127 bytes back), plus their plain timing.
TAG=${1:-r08b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in default near near64; do
  LIB=""
  if [ "$v" != "default" ]; then LIB=$PWD/rust-brotli-decompressor_b200/variants/libbrotli_b200_$v.so; fi
  if [ $v = near64 ]; then
    BROTLI_B200_LIB=$LIB timeout 300 python bench.py --streams 189440 --unique 2048 --steps 3 --warmup 3 --no-e2e --no-cpu --no-other-configs > $OUT/bench_$v.json 2> $OUT/bench_$v.err
    python -c "import json; j=json.load(open('$OUT/bench_$v.json')); print('$v headline probe', j['value'], 'GB/s ms', j['ms_per_step'], 'bit_exact', j.get('bit_exact'))"
  fi
  BROTLI_B200_LIB=$LIB timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum \
    --clock-control none -k regex:brotli_decode_lane -s 3 -c 1 --csv --log-file $OUT/traffic_$v.csv \
    python bench.py --streams 94720 --unique 2048 --steps 1 --warmup 3 --no-e2e --no-cpu --no-other-configs > $OUT/traffic_bench_$v.log 2>&1
  echo "== $v"; grep -v "^==" $OUT/traffic_$v.csv | cut -d, -f5,13- | tail -8
done 2>&1 | tee $OUT/summary.txt

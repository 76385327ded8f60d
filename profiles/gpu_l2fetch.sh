#!/bin/bash
# One gpurun call: device-resident headline probe + DRAM traffic of the lane kernel for L2 fetch granularities.
TAG=${1:-l2fetch}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for f in 0 32 64 128; do
  export BROTLI_B200_L2_FETCH=$f
  timeout 600 python bench.py --streams 131072 --unique 2048 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/bench_f$f.json 2> $OUT/bench_f$f.err
  python -c "import json; j=json.load(open('$OUT/bench_f$f.json')); print('fetch $f value', j['value'], 'GB/s ms', j['ms_per_step'], 'bit_exact', j.get('bit_exact'))"
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:brotli_decode_lane -s 3 -c 1 --csv --log-file $OUT/traffic_f$f.csv \
    python bench.py --streams 131072 --unique 2048 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/traffic_bench_f$f.log 2>&1
  grep -v "^==" $OUT/traffic_f$f.csv | cut -d, -f13- | tail -3
done

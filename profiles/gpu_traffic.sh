#!/bin/bash
# One gpurun call: device-resident headline probe (131 072 streams) plus the DRAM traffic of one lane-kernel launch.
#   gpurun --timeout 900 -- 'bash profiles/gpu_traffic.sh r04c "14"'
TAG=${1:-traffic}
WARPS=${2:-"14"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for w in $WARPS; do
  BROTLI_B200_LANE_WARPS=$w timeout 600 python bench.py --streams 131072 --unique 2048 --steps 3 --warmup 3 --no-e2e --no-cpu > $OUT/bench_w$w.json 2> $OUT/bench_w$w.err
  python -c "import json; j=json.load(open('$OUT/bench_w$w.json')); print('warps $w value', j['value'], 'GB/s ms', j['ms_per_step'], 'bit_exact', j.get('bit_exact'))"
  BROTLI_B200_LANE_WARPS=$w timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,lts__t_sector_hit_rate.pct \
    --clock-control none -k regex:brotli_decode_lane -s 3 -c 1 --csv --log-file $OUT/traffic_w$w.csv \
    python bench.py --streams 131072 --unique 2048 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/traffic_bench_w$w.log 2>&1
  grep -v "^==" $OUT/traffic_w$w.csv | cut -d, -f5,13- | tail -7
done

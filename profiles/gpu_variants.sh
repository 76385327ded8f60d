#!/bin/bash
# One gpurun call: headline bench (device-resident leg only) for library variants x lane-kernel geometries.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_variants.sh r01j "default lit2" "8 16"'
TAG=${1:-var}
VARS=${2:-"default"}
WARPS=${3:-"8 16"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in $VARS; do
  LIB=""
  if [ "$v" != "default" ]; then LIB=$PWD/rust-brotli-decompressor_b200/variants/libbrotli_b200_$v.so; fi
  for w in $WARPS; do
    BROTLI_B200_LIB=$LIB BROTLI_B200_LANE_WARPS=$w timeout 600 python bench.py --streams ${STREAMS:-131072} --unique 2048 --steps 3 --warmup 3 --no-e2e --no-cpu \
      > $OUT/bench_${v}_w$w.json 2> $OUT/bench_${v}_w$w.err
    python - <<PY
import json
try:
    j = json.load(open("$OUT/bench_${v}_w$w.json"))
    print("$v warps $w value", j["value"], "GB/s ms", j["ms_per_step"], "bit_exact", j.get("bit_exact"))
except Exception as e:
    print("$v warps $w: no result", e); print(open("$OUT/bench_${v}_w$w.err").read()[-1500:])
PY
  done
done
if [ -n "$PROF_WARPS" ]; then
  BROTLI_B200_LANE_WARPS=$PROF_WARPS timeout 900 ncu --set full --clock-control none --import-source on -k regex:brotli_decode_lane -s 3 -c 1 -o $OUT/prof_lane \
    python bench.py --streams 131072 --unique 2048 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_bench_lane.log 2>&1
  tail -2 $OUT/prof_bench_lane.log
fi

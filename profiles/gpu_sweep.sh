#!/bin/bash
# One gpurun call: parity tests, then the headline bench (device-resident leg only) for several
# lane-kernel geometries, then one full ncu capture of the lane kernel.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_sweep.sh r01d "4 6 8 12 16"'
TAG=${1:-sweep}
WARPS=${2:-"4 8"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
for w in $WARPS; do
  BROTLI_B200_LANE_WARPS=$w timeout 600 python bench.py --streams ${STREAMS:-131072} --unique 2048 --steps 3 --warmup 3 --no-e2e --no-cpu \
    > $OUT/bench_w$w.json 2> $OUT/bench_w$w.err
  echo "warps $w exit $?"; python - <<PY
import json
try:
    j = json.load(open("$OUT/bench_w$w.json"))
    print("warps $w value", j["value"], "GB/s ms", j["ms_per_step"], "bit_exact", j.get("bit_exact"), "roofline", j["roofline"]["achieved"])
except Exception as e:
    print("no result", e); print(open("$OUT/bench_w$w.err").read()[-2000:])
PY
done
if [ -n "$PROF_WARPS" ]; then
  BROTLI_B200_LANE_WARPS=$PROF_WARPS timeout 900 ncu --set full --clock-control none --import-source on -k regex:brotli_decode_lane -s 3 -c 1 -o $OUT/prof_lane \
    python bench.py --streams 131072 --unique 2048 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_bench_lane.log 2>&1
  tail -2 $OUT/prof_bench_lane.log
fi
ls -la $OUT

#!/bin/bash
# One gpurun call: parity tests, bench line (both arms), ncu launch list and one full ncu capture of the
# dominant decode kernel.  Everything lands in gpurun_out/<tag>/ ; summaries are copied to profiles/ by hand.
#   gpurun --timeout 1700 -- 'bash profiles/gpu_round.sh r01z'
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt; free -g >> $OUT/gpu.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} > $OUT/bench.json 2> $OUT/bench.err
echo "bench exit $?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
if [ -n "$WITH_REFERENCE" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
  cat $OUT/bench_reference.json
fi
# launch list (cold-cache, serialised: compare shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches.csv \
  python bench.py --streams 131072 --unique 2048 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_bench.log 2>&1
# full capture of the dominant decode kernel (4th launch = first timed step)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:brotli_decode_lane -s 3 -c 1 -o $OUT/prof_lane \
  python bench.py --streams 131072 --unique 2048 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_bench.log 2>&1
ls -la $OUT

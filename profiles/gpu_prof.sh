#!/bin/bash
# Profile-only gpurun job: full ncu capture of one kernel (regex $2, default team) on a 32 Ki-stream launch.
#   gpurun --timeout 900 -- 'bash profiles/gpu_prof.sh r01c team'
TAG=${1:-prof}
KERN=${2:-team}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 800 ncu --set full --clock-control none --import-source on -k regex:$KERN -s 3 -c 1 -o $OUT/prof_$KERN \
  python bench.py --streams 32768 --unique 1024 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_bench_$KERN.log 2>&1
tail -2 $OUT/prof_bench_$KERN.log
ls -la $OUT

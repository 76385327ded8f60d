#!/bin/bash
# One gpurun call: source-level ncu captures of the lane kernel on other configs (C3, C5, ...), to see where their
# rounds and headers go.  gpurun --timeout 1500 -- 'bash profiles/gpu_prof_cfg.sh r05p "C3 C5"'
TAG=${1:-prof}
CFGS=${2:-"C3 C5"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for cfg in $CFGS; do
  n=131072; if [ $cfg = C3 ]; then n=524288; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:brotli_decode_lane -s 2 -c 1 -o $OUT/prof_$cfg \
    python bench.py --config $cfg --streams $n --unique 2048 --steps 1 --warmup 2 --no-e2e --no-cpu --no-other-configs > $OUT/prof_$cfg.log 2>&1
  python profiles/ncu_hot.py $OUT/prof_$cfg.ncu-rep > $OUT/ncu_$cfg.txt 2>&1
  head -20 $OUT/ncu_$cfg.txt
done
ls -la $OUT

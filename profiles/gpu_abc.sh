#!/bin/bash
# One gpurun call: A/B of library variants on the headline probe (two repetitions) and on C3 / C5 (once), same box.
#   gpurun --timeout 1200 -- 'bash profiles/gpu_abc.sh <tag> "default v1 v2"'
TAG=${1:-abc}
VARS=${2:-"default"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for rep in 1 2; do
for v in $VARS; do
  LIB=""
  if [ "$v" != "default" ]; then LIB=$PWD/rust-brotli-decompressor_b200/variants/libbrotli_b200_$v.so; fi
  BROTLI_B200_LIB=$LIB timeout 300 python bench.py --streams ${STREAMS:-189440} --unique 2048 --steps 3 --warmup 3 --no-e2e --no-cpu --no-other-configs > $OUT/bench_${v}_$rep.json 2> $OUT/bench_${v}_$rep.err
  python -c "import json; j=json.load(open('$OUT/bench_${v}_$rep.json')); print('$v rep $rep headline probe', j['value'], 'GB/s ms', j['ms_per_step'], 'bit_exact', j.get('bit_exact'))"
  if [ $rep = 1 ]; then
    CONFIGS=${CONFIGS:-C3,C5} BROTLI_B200_LIB=$LIB timeout 300 python profiles/gpu_configs.py > $OUT/configs_$v.jsonl 2> $OUT/configs_$v.err
    python -c "
import json
for l in open('$OUT/configs_$v.jsonl'):
    j=json.loads(l); print('$v', j['config'], j['GBps'], 'GB/s', j['ms'], 'ms bailed', j['bailed_to_exact'], j['bit_exact'])"
  fi
done
done 2>&1 | tee $OUT/summary.txt

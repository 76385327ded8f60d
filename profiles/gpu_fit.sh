#!/bin/bash
# One gpurun call: geometry by wave fit (decided on the device for uniform batches) on / off for the batch sizes a
# multi-GPU run hands one GPU, then the other configs (which must not be fitted: their streams are not uniform).
TAG=${1:-fit}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for n in 262144 131072 65536; do
  for fit in 1 0; do
    BROTLI_B200_LANE_FIT=$fit timeout 600 python bench.py --streams $n --steps 3 --warmup 2 --unique 2048 --no-e2e --no-cpu --no-other-configs > $OUT/bench_n${n}_fit$fit.json 2> $OUT/bench_n${n}_fit$fit.err
    python -c "import json; j=json.load(open('$OUT/bench_n${n}_fit$fit.json')); print('n=$n fit=$fit', j['value'], 'GB/s ms', j['ms_per_step'], 'bit_exact', j.get('bit_exact'))"
  done
done
timeout 900 python profiles/gpu_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err
python -c "
import json
for l in open('$OUT/configs.jsonl'):
    j=json.loads(l); print(j['config'], j['GBps'], 'GB/s', j['ms'], 'ms bailed', j['bailed_to_exact'], j['bit_exact'])"

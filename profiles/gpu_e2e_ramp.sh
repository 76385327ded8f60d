#!/bin/bash
# One gpurun call: chunk plans of the host-buffer pipeline (BROTLI_B200_PIPE_RAMP) on the e2e leg of the headline bench.
TAG=${1:-ramp}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for r in "" 1 2 3; do
  BROTLI_B200_PIPE_RAMP=$r timeout 600 python bench.py --steps 3 --warmup 2 --unique 2048 --no-cpu --no-other-configs > $OUT/bench_ramp$r.json 2> $OUT/bench_ramp$r.err
  python -c "import json; j=json.load(open('$OUT/bench_ramp$r.json')); e=j['e2e']; print('ramp=$r value', j['value'], 'e2e', e['value'], 'GB/s', e['ms_per_step'], 'ms copy-only', e.get('copy_only_ms_per_step'), 'bit_exact', e['bit_exact'])"
done

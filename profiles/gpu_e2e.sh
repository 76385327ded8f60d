#!/bin/bash
# One gpurun call: the host-buffer (e2e) leg of the headline bench for pipeline ramp settings.
TAG=${1:-e2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for r in ${2:-"3 2 4"}; do
  BROTLI_B200_PIPE_RAMP=$r timeout 800 python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/bench_ramp$r.json 2> $OUT/bench_ramp$r.err
  python -c "import json; j=json.load(open('$OUT/bench_ramp$r.json')); print('ramp $r value', j['value'], 'e2e', j['e2e']['value'], 'GB/s ms', j['e2e']['ms_per_step'], 'span', j['e2e']['last_kernel_span_ms'], 'exact', j['e2e']['bit_exact'])"
done

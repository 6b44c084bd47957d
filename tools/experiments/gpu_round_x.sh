#!/bin/bash
mkdir -p gpurun_out
for cfg in "131072 16384" "131072 65536" "262144 32768" "65536 16384" "196608 49152"; do
  set -- $cfg
  AFC_CHUNK_ITEMS=$1 AFC_MIN_CHUNK_ITEMS=$2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print('$cfg', 'e2e ms', round(d['e2e']['ms_per_step'],3), 'value ms', round(d['ms_per_step'],3))"
done

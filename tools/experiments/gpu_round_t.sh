#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "canonical or config1" > gpurun_out/pytest_canon.log 2>&1; tail -3 gpurun_out/pytest_canon.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_t.json')); x=d['extras']['canonical_form_vc_documents']
for k in ('plain_values','6pct_escaped'): print(k, round(x[k]['ms'],3), {a:round(b,3) for a,b in x[k]['kernels_ms'].items() if 'json' in a})"

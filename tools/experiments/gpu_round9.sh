#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras > gpurun_out/bench_r9.json 2> gpurun_out/bench_r9.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r9.json')); e=d['extras']
print('value',d['value'],'e2e',d['e2e']['value'],'hram',d['roofline']['other_kernels_ms'],'verify ms',d['roofline']['kernel_avg_ms']); print('keyed',d['keyed']['value'])
for k,v in e.items():
    if k not in ('microbench','ingest_soak_100k','ingest_soak_1M'): print(k, v)
print(e['microbench']['sha256_compress'])"

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras > gpurun_out/bench_keyed.json 2> gpurun_out/bench_keyed.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_keyed.err
python -c "
import json; d=json.load(open('gpurun_out/bench_keyed.json')); print('value',d['value'],'e2e',d['e2e']['value']); print('keyed',d['keyed']); e=d['extras']
for k,v in e.items():
    if k!='microbench': print(k, v)
for k in ('sha256_compress','sha512_compress'): print(k, e['microbench'][k])"

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r12.json 2> gpurun_out/bench_r12.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r12.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r12.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches']); print('roofline',d['roofline']); print('keycache',d['keycache']); print('no_keycache',d['no_keycache']); print('keyed',d['keyed'])"

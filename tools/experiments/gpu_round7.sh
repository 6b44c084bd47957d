#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras > gpurun_out/bench_r7.json 2> gpurun_out/bench_r7.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r7.json')); e=d['extras']
print('value',d['value'],'e2e',d['e2e']['value'],'hram',d['roofline']['other_kernels_ms'],'verify ms',d['roofline']['kernel_avg_ms']); print('keyed',d['keyed'])
for k,v in e.items():
    if k!='microbench': print(k, v)"
# memory-safety evidence: compute-sanitizer memcheck over the small parity tests (golden vectors, ragged lengths, empty batches, Merkle appends)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
   -k "golden or empty or ragged_lengths or incremental or codecs or expanded_key_cache_sign" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -15 gpurun_out/sanitizer_memcheck.log

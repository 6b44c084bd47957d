#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_w.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'warm', d['warm_keycache']['ms_per_step'], 'keyed', d['keyed']['ms_per_step'])"
timeout 300 compute-sanitizer --tool initcheck --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "golden or ragged or empty or incremental or audit_proofs" > gpurun_out/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -2 gpurun_out/sanitizer_initcheck.log

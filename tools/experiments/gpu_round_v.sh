#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['warm_keycache']['value'])"

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r.json')); x=d['extras']; print(d['value'], d['ms_per_step']); print(x.get("sign_512B_expanded_keys"), x.get("sign_512B_from_seeds")); print(x.get('merkle_append_64B_leaves'))"
tail -3 gpurun_out/bench_r.err

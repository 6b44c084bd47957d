#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none -k regex:"^k_(ed_|hmac|sha256|merkle_leaf|merkle_level)" -s 17 -c 36 -o /tmp/prof_all -f python tools/ncu_targets.py > gpurun_out/ncu_all.log 2>&1
tail -3 gpurun_out/ncu_all.log
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2>/dev/null
ls -la /tmp/prof_all.ncu-rep gpurun_out/prof_all_raw.csv
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras > gpurun_out/bench_soak.json 2> gpurun_out/bench_soak.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_soak.json')); e=d['extras']; print(e.get('ingest_soak_100k')); print(e.get('ingest_soak_1M'))"

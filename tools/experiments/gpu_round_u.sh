#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
VARIANTS="_load32" bash tools/experiments/gpu_round_p.sh
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err
python -c "
import json; x=json.load(open('gpurun_out/bench_u.json'))['extras']; print({k:round(x[k]['ms'],3) for k in ('sign_512B_expanded_keys','sign_512B_from_seeds')})"

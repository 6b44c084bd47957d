#!/bin/bash
# First GPU pass: parity tests, smoke, bench (+extras), ncu launch list, ncu full capture of the verify kernels.
set -x
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --extras > gpurun_out/bench_extras.json 2> gpurun_out/bench_extras.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/bench_extras.json; tail -5 gpurun_out/bench_extras.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 1500 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
tail -20 gpurun_out/launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_ed_verify|k_ed_hram" -s 2 -c 2 -o gpurun_out/prof_verify -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out

#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_multi.log
tail -15 gpurun_out/pytest_gpu_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --extras > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
tail -c 3000 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "ref rc=$?"
tail -c 600 gpurun_out/bench_2gpu_ref.json
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -c 2500 gpurun_out/bench_1gpu.json

#!/bin/bash
# round 2, run d (2 GPUs): all gpu tests incl. the NCCL ones, bench at N = 2 (cfg4 with the all-gather), soak at N = 2, reference arm
mkdir -p gpurun_out/r2d
nvidia-smi topo -m > gpurun_out/r2d/topo.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2d/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest_gpu.log
tail -8 gpurun_out/r2d/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2d/bench_2gpu.json 2> gpurun_out/r2d/bench_2gpu.err; echo "bench 2gpu rc=$?"
grep -c "NCCL INFO" gpurun_out/r2d/bench_2gpu.err; grep -E "nranks|Init COMPLETE|ncclCommInitRank" gpurun_out/r2d/bench_2gpu.err | head -6
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2d/bench_2gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e', {k:v for k,v in d['e2e'].items() if k!='note'})
    print('cfg4', d['cfg4']); print('warm', d.get('warm_keycache',{}).get('ms_per_step'))
except Exception as e:
    print('parse fail', e); print(open('gpurun_out/r2d/bench_2gpu.err').read()[-3000:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --soak 10 > gpurun_out/r2d/soak_2gpu.json 2> gpurun_out/r2d/soak_2gpu.err; echo "soak 2gpu rc=$?"
tail -1 gpurun_out/r2d/soak_2gpu.json | cut -c1-900
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2d/ref_2gpu.json 2> gpurun_out/r2d/ref_2gpu.err; echo "ref rc=$?"; tail -1 gpurun_out/r2d/ref_2gpu.json | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 50 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "bench$n rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/bench_${n}gpu.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print({k:d[k] for k in ("value","n_gpus","ms_per_step","clocks","gpu_launches")}, "e2e", d["e2e"]["value"], "warm", d["warm_keycache"].get("value"), "nocache", d["no_keycache"].get("value"))
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29630 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_8gpu_ref.json 2> gpurun_out/bench_8gpu_ref.err; echo "ref rc=$?"
tail -c 400 gpurun_out/bench_8gpu_ref.json
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_multi.log 2>&1; tail -3 gpurun_out/pytest_gpu_multi.log

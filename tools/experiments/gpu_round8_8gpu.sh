#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "bench$n rc=$?"
tail -2 gpurun_out/bench_${n}gpu.err
python - <<PY
import json
for ln in open("gpurun_out/bench_${n}gpu.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print({k:d[k] for k in ("value","n_gpus","ms_per_step","clocks","e2e","gpu_launches")}); print(d.get("keyed"))
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 3 --warmup 3 --extras --no-cpu-baseline > gpurun_out/bench_8gpu_extras.json 2> gpurun_out/bench_8gpu_extras.err; echo "extras rc=$?"
python - <<PY
import json
for ln in open("gpurun_out/bench_8gpu_extras.json"):
    if ln.startswith("{"):
        d=json.loads(ln); e=d["extras"]; print({k:v for k,v in e.items() if k!="microbench"})
PY

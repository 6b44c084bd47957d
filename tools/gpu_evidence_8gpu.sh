#!/bin/bash
# round 2, 8 GPUs: bench (cfg4 with the NCCL all-gather over 8 ranks, NUMA-bound ranks), 60 s soak at 100 k actions/s over 8 GPUs, reference arm
out=gpurun_out/r2_8gpu; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 3 > $out/bench_8gpu.json 2> $out/bench_8gpu.err; echo "bench 8gpu rc=$?"
grep -c "NCCL INFO" $out/bench_8gpu.err; grep -E "Init COMPLETE" $out/bench_8gpu.err | head -3 | cut -c1-220
python - <<'PY'
import json
o='gpurun_out/r2_8gpu/'
try:
    d=json.loads(open(o+'bench_8gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e', {k:v for k,v in d['e2e'].items() if k!='note'})
    print('cfg4', d['cfg4']); print('warm', d.get('warm_keycache',{}).get('value'), 'nocache', d.get('no_keycache',{}).get('value'))
except Exception as e:
    print('parse fail', e); print(open(o+'bench_8gpu.err').read()[-3000:])
PY
[ -n "$SKIP_SOAK" ] || timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --soak 60 > $out/soak_8gpu_60s.json 2> $out/soak_8gpu_60s.err; echo "soak 8gpu rc=$?"
[ -n "$SKIP_SOAK" ] || tail -1 $out/soak_8gpu_60s.json | cut -c1-1600
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $out/ref_8gpu.json 2> $out/ref_8gpu.err; echo "ref rc=$?"; tail -1 $out/ref_8gpu.json | cut -c1-200
ls -la $out

#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d=json.loads(ln)
        print(sys.argv[1], {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'clocks', d['clocks'])
        print('  e2e', {k:v for k,v in d['e2e'].items() if k!='note'}); print('  warm', d.get('warm_keycache',{}).get('ms_per_step'), 'nocache', d.get('no_keycache',{}).get('ms_per_step'), 'keyed', d.get('keyed',{}).get('ms_per_step'))
        print('  kernels', d['roofline']['kernel'], d['roofline']['kernel_avg_ms'], {k:round(v,4) for k,v in d['roofline']['other_kernels_ms'].items()})
PY
}
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; echo "bench rc=$?"; summ gpurun_out/bench_m.json

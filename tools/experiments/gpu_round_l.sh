#!/bin/bash
# round-1 experiment L: radix-65536 base table, run-time group size, hram on a side stream, 4-slot host pipeline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d=json.loads(ln)
        print(sys.argv[1], {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'clocks', d['clocks'])
        print('  e2e', d['e2e']); print('  warm', d.get('warm_keycache',{}).get('ms_per_step'), 'nocache', d.get('no_keycache',{}).get('ms_per_step'), 'keyed', d.get('keyed',{}).get('ms_per_step'))
        print('  kernels', d['roofline']['kernel'], d['roofline']['kernel_avg_ms'], {k:round(v,4) for k,v in d['roofline']['other_kernels_ms'].items()})
        x=d.get('extras')
        if x: print('  extras', json.dumps(x)[:1500])
PY
}
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --extras > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "new rc=$?"; summ gpurun_out/bench_new.json
AFC_KC_GROUP=4 timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g4.json 2> gpurun_out/bench_g4.err; echo "g4 rc=$?"; summ gpurun_out/bench_g4.json
AFC_LIB=$PWD/build/ab/libafcrypto_w8.so timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w8.json 2> gpurun_out/bench_w8.err; echo "w8 rc=$?"; summ gpurun_out/bench_w8.json
for g in 5 6 7 8; do AFC_KC_GROUP=$g timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$g.json 2> gpurun_out/bench_g$g.err; summ gpurun_out/bench_g$g.json | head -3; done

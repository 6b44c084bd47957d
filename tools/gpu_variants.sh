#!/bin/bash
mkdir -p gpurun_out/r2v
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2v/bench_$name.json 2> gpurun_out/r2v/bench_$name.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2v/bench_$name.json')); r=d['roofline']; o=r['other_kernels_ms']
    print('$name', 'cold %.3f'%d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'verify %.3f'%r['kernel_avg_ms'], 'rows %.3f hram %.3f'%(o.get('k_kc_rows',0), o.get('k_ed_hram',0)), 'cfg4 %.2f fast %.2f'%(d['cfg4']['ms'], d['cfg4']['fast_variable_time_ms']))
except Exception as e: print('$name parse fail rc=$rc', e); print(open('gpurun_out/r2v/bench_$name.err').read()[-1500:])
PY
}
run main
for v in fold t64 minb4 nopf pfl2 g8; do run $v AFC_LIB=$PWD/agentfield_b200/variants/libafcrypto_$v.so; done
AFC_LIB=$PWD/agentfield_b200/variants/libafcrypto_fold.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "selftest or golden or random_parity or transparent or keyed" 2>&1 | tail -3

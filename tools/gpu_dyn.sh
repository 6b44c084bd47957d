#!/bin/bash
# counter-driven vs static split of the table-driven verify kernels: parity tests, then bench.py --ab for each
mkdir -p gpurun_out/r2y
[ -n "$SKIP_TESTS" ] || timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --ab > gpurun_out/r2y/bench_$name.json 2> gpurun_out/r2y/bench_$name.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2y/bench_$name.json')); r=d['roofline']; o=r['other_kernels_ms']
    print('$name', 'cold %.3f'%d['ms_per_step'], 'warm %.3f'%d['warm_keycache']['ms_per_step'], 'keyed %.3f'%d['keyed']['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), r['kernel'], '%.3f'%r['kernel_avg_ms'], {k: round(v,3) for k,v in o.items() if v > 0.05})
except Exception as e: print('$name parse fail rc=$rc', e); print(open('gpurun_out/r2y/bench_$name.err').read()[-1500:])
PY
}
run dyn1 AFC_VERIFY_DYNAMIC=1
[ -n "$SKIP_DYN0" ] || run dyn0 AFC_VERIFY_DYNAMIC=0
for v in $VARIANTS; do run $v AFC_VERIFY_DYNAMIC=1 AFC_LIB=$PWD/agentfield_b200/variants/libafcrypto_$v.so; done

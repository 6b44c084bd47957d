#!/bin/bash
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
for ln in open(sys.argv[1]):
    if ln.startswith("{"):
        d=json.loads(ln)
        print(sys.argv[1], {k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['ms_per_step'],3), 'warm', d.get('warm_keycache',{}).get('ms_per_step'), 'keyed', d.get('keyed',{}).get('ms_per_step'))
        print('  kernels', d['roofline']['kernel'], d['roofline']['kernel_avg_ms'], {k:round(v,4) for k,v in d['roofline']['other_kernels_ms'].items()})
PY
}
for v in "" $VARIANTS; do
  lib=$PWD/agentfield_b200/libafcrypto.so; [ -n "$v" ] && lib=$PWD/build/ab/libafcrypto$v.so
  AFC_LIB=$lib timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p$v.json 2> gpurun_out/bench_p$v.err; echo "bench$v rc=$?"; summ gpurun_out/bench_p$v.json
done

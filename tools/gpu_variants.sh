#!/bin/bash
# A/B of compile-time variants of the table-driven verify kernel.  Build them HERE first (nvcc cross-compiles):
#   bash tools/build_variant.sh fold  "-DAFC_FOLD_SHIFT=1"
#   bash tools/build_variant.sh t64   "-DAFC_CACHED_THREADS=64 -DAFC_CACHED_MINB=6"
#   bash tools/build_variant.sh minb4 "-DAFC_CACHED_MINB=4"
#   bash tools/build_variant.sh nopf  "-DAFC_KP_PREFETCH=0"
#   bash tools/build_variant.sh pfl2  "-DAFC_KP_PREFETCH=2"
#   bash tools/build_variant.sh g16   "-DAFC_KC_GMAX=16"
# then: gpurun -- 'VARIANTS="fold t64 minb4 nopf pfl2 g16" bash tools/gpu_variants.sh'   (variants travel with the snapshot, git ignores them)
mkdir -p gpurun_out/r2v
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --ab > gpurun_out/r2v/bench_$name.json 2> gpurun_out/r2v/bench_$name.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2v/bench_$name.json')); r=d['roofline']; o=r['other_kernels_ms']
    print('$name', 'cold %.3f'%d['ms_per_step'], 'warm %.3f'%d['warm_keycache']['ms_per_step'], 'keyed %.3f'%d['keyed']['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), r['kernel'], '%.3f'%r['kernel_avg_ms'], 'rows %.3f hram %.3f'%(o.get('k_kc_rows',0), o.get('k_ed_hram',0)))
except Exception as e: print('$name parse fail rc=$rc', e); print(open('gpurun_out/r2v/bench_$name.err').read()[-1500:])
PY
}
run main
for v in $VARIANTS; do run $v AFC_LIB=$PWD/agentfield_b200/variants/libafcrypto_$v.so; done

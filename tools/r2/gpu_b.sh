#!/bin/bash
# round 2, run b: cooperative chain (4 lanes per key) on/off, occupancy variants of k_kc_rows, new parity tests
mkdir -p gpurun_out/r2b
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2b/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest_gpu.log
tail -15 gpurun_out/r2b/pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2b/bench_$name.json 2> gpurun_out/r2b/bench_$name.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2b/bench_$name.json'))
    r=d['roofline']; o=r['other_kernels_ms']
    print('$name', 'cold %.3f'%d['ms_per_step'], 'warm %.3f'%d['warm_keycache'].get('ms_per_step',0), 'e2e %.1fM'%(d['e2e']['value']/1e6), 'verify %.3f'%r['kernel_avg_ms'], {k:round(v,3) for k,v in o.items() if k in ('k_kc_chain','k_kc_chain4','k_kc_rows','k_ed_hram','k_kc_scatter','k_kc_dedup')})
except Exception as e: print('$name parse fail rc=$rc', e); print(open('gpurun_out/r2b/bench_$name.err').read()[-1500:])
PY
}
run main_chain4 AFC_KC_CHAIN4=1
run main_chain1 AFC_KC_CHAIN4=0
for v in r256_3 r256_4 r128_4 r128_6 r128_8; do
  [ -f agentfield_b200/variants/libafcrypto_$v.so ] && run $v AFC_LIB=$PWD/agentfield_b200/variants/libafcrypto_$v.so
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"^k_(kc_chain4|kc_rows|ed_verify_cached|ed_hram)" -c 8 -o gpurun_out/r2b/prof_cached -f python tools/r2/ncu_cached.py > gpurun_out/r2b/ncu.log 2>&1
tail -2 gpurun_out/r2b/ncu.log
ncu -i gpurun_out/r2b/prof_cached.ncu-rep --page raw --csv > gpurun_out/r2b/prof_cached_raw.csv 2>/dev/null
rm -f gpurun_out/r2b/prof_cached.ncu-rep
ls -la gpurun_out/r2b/

#!/bin/bash
# round 2, run a: new issuer-key cache (bucketed order, hot/cold, LRU, staged table build) — tests, bench, stage variants, ncu
mkdir -p gpurun_out/r2a
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2a/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest_gpu.log
tail -15 gpurun_out/r2a/pytest_gpu.log
for st in 4 1 2; do
  AFC_KC_STAGES=$st timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/r2a/bench_st$st.json 2> gpurun_out/r2a/bench_st$st.err; echo "bench stages=$st rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2a/bench_st$st.json'))
    print('stages $st', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'warm', d['warm_keycache'].get('ms_per_step'), 'nocache', d['no_keycache'].get('ms_per_step'), 'keyed', d['keyed'].get('ms_per_step'))
    r=d['roofline']; print('  dom', r['kernel'], r['kernel_avg_ms'], {k:round(v,4) for k,v in r['other_kernels_ms'].items()})
except Exception as e: print('parse fail', e); print(open('gpurun_out/r2a/bench_st$st.err').read()[-2000:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"^k_(ed_verify_cached|ed_hram|kc_)" -c 40 -o gpurun_out/r2a/prof_cached -f python tools/r2/ncu_cached.py > gpurun_out/r2a/ncu.log 2>&1
tail -3 gpurun_out/r2a/ncu.log
ncu -i gpurun_out/r2a/prof_cached.ncu-rep --page raw --csv > gpurun_out/r2a/prof_cached_raw.csv 2>/dev/null; ls -la gpurun_out/r2a/

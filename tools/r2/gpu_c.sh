#!/bin/bash
# round 2, run c: new bench.py (cfg4 / cfg3 / issuer mix / NUMA in the default line), G <= 16 vs G <= 8, soak mode
mkdir -p gpurun_out/r2c
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2c/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest_gpu.log
tail -5 gpurun_out/r2c/pytest_gpu.log
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/r2c/bench_default.json 2> gpurun_out/r2c/bench_default.err; echo "bench default rc=$? in $(( $(date +%s) - t0 )) s"
tail -3 gpurun_out/r2c/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['clocks'])
print('e2e', {k:v for k,v in d['e2e'].items() if k not in('note',)})
r=d['roofline']; print('roofline', {k:r[k] for k in ('kernel','frac','frac_step','kernel_avg_ms','kernel_share_of_step','kernel_launches_timed','untimed_launches','traffic')})
print('others', {k:round(v,3) for k,v in r['other_kernels_ms'].items()})
for k in ('cfg4','cfg3','warm_keycache','no_keycache','keyed','issuer_mix','keycache','cpu_baseline'):
    print(k, json.dumps(d.get(k))[:600])
PY
for v in main g8; do
  if [ $v = main ]; then E=""; else E="AFC_LIB=$PWD/agentfield_b200/variants/libafcrypto_$v.so"; fi
  env $E timeout 300 python bench.py --steps 30 --no-secondary --no-cpu-baseline > gpurun_out/r2c/bench_$v.json 2> gpurun_out/r2c/bench_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2c/bench_$v.json')); r=d['roofline']; print('$v cold %.3f verify %.3f e2e %.1fM cfg4 %s'%(d['ms_per_step'], r['kernel_avg_ms'], d['e2e']['value']/1e6, {k:d['cfg4'].get(k) for k in ('ms','root_ok')}))"
done
timeout 300 python bench.py --soak 10 > gpurun_out/r2c/soak.json 2> gpurun_out/r2c/soak.err; echo "soak rc=$?"; cut -c1-1500 gpurun_out/r2c/soak.json; tail -3 gpurun_out/r2c/soak.err

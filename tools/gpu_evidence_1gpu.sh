#!/bin/bash
# round-2 single-GPU evidence: tests, smoke, the bench arms, soak, launch list, ncu --set full of every hot kernel
out=gpurun_out/r2final; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" >> $out/smoke.log; tail -2 $out/smoke.log
./tests/c_abi/harness > $out/c_abi_harness.log 2>&1; echo "harness rc=$?" >> $out/c_abi_harness.log; tail -2 $out/c_abi_harness.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --extras > $out/bench_extras.json 2> $out/bench_extras.err; echo "extras rc=$?"
timeout 300 python bench.py --soak 60 > $out/soak_60s.json 2> $out/soak_60s.err; echo "soak rc=$?"
python - <<'PY'
import json
o='gpurun_out/r2final/'
d=json.load(open(o+'bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches','clocks')}); print('e2e',{k:v for k,v in d['e2e'].items() if k!='note'})
r=d['roofline']; print('roofline',{k:v for k,v in r.items() if k not in ('note','other_kernels_ms')}); print({k:round(v,3) for k,v in r['other_kernels_ms'].items()})
for k in ('cfg4','cfg3','warm_keycache','no_keycache','keyed','issuer_mix'): print(k, json.dumps(d.get(k))[:500])
print('cpu',d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['openssl'].get('value'))
r=json.load(open(o+'bench_reference.json')); print('reference arm',r['value'],r['cpu_baseline']['cores'], r['config']==d['config'])
x=json.load(open(o+'bench_extras.json'))['extras']; print({k:(v.get('ms'), v.get('signs_per_s')) for k,v in x.items() if k!='microbench' and isinstance(v,dict)})
s=json.load(open(o+'soak_60s.json')); print('soak', s['value'], s['latency_us'], s['pushed'], s['cpu_baseline']['value'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 40 -c 80 --csv --log-file $out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $out/ncu_launch.log 2>&1
timeout 1800 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"^k_(ed_|hmac|sha256|merkle_leaf|merkle_level|kc_|json)" -c 90 -o $out/prof_all -f python tools/ncu_targets.py > $out/ncu_all.log 2>&1
tail -2 $out/ncu_all.log
ncu -i $out/prof_all.ncu-rep --page raw --csv > $out/prof_all_raw.csv 2>/dev/null
ncu -i $out/prof_all.ncu-rep --page source --csv --kernel-name regex:k_ed_verify_cached --print-source sass > /tmp/verify_cached_source.csv 2>/dev/null
python tools/ncu_stalls.py /tmp/verify_cached_source.csv > $out/verify_cached_stalls.txt 2>&1
python tools/ncu_summary.py $out/prof_all_raw.csv > $out/ncu_all_kernels_summary.txt 2>&1
rm -f $out/prof_all.ncu-rep        # 120 MB: gpurun_out/ is capped at 64 MiB
ls -la $out/

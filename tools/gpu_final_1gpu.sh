#!/bin/bash
# round-end single-GPU evidence: tests, smoke, the three bench arms, launch list, ncu --set full of the default path and of every kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --extras > gpurun_out/bench_extras.json 2> gpurun_out/bench_extras.err; echo "extras rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches','clocks')}); print('e2e',{k:v for k,v in d['e2e'].items() if k!='note'}); print('roofline',{k:v for k,v in d['roofline'].items() if k!='note'}); print('warm',d['warm_keycache']['ms_per_step']); print('nocache',d['no_keycache']['ms_per_step']); print('keyed',d['keyed']['ms_per_step']); print('cpu',d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
r=json.load(open('gpurun_out/bench_reference.json')); print('reference arm',r['value'],r['cpu_baseline']['cores'])
x=json.load(open('gpurun_out/bench_extras.json'))['extras']; print({k:(v.get('ms'), v.get('p99_us')) for k,v in x.items() if k!='microbench'})"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 33 -c 66 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none --profile-from-start off -k regex:"^k_(ed_|hmac|sha256|merkle_leaf|merkle_level|kc_build|kc_bases|json)" -c 64 -o /tmp/prof_all -f python tools/ncu_targets.py > gpurun_out/ncu_all.log 2>&1
tail -2 gpurun_out/ncu_all.log
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2>/dev/null; ls -la gpurun_out/prof_all_raw.csv gpurun_out/launches_final.csv

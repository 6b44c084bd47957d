#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --extras > gpurun_out/bench_r14.json 2> gpurun_out/bench_r14.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r14.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r14.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks']); print('roofline',d['roofline']); print('keycache',d['keycache']); print('no_keycache',d['no_keycache']); print('keyed',d['keyed']); print('cpu',d['cpu_baseline'])
e=d['extras']
for k,v in e.items():
    if k!='microbench': print(k,v)"
timeout 900 ncu --set full --clock-control none -k regex:"^k_(ed_verify_cached|ed_hram|kc_dedup|kc_build)" -s 4 -c 4 -o /tmp/prof_cached -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cached.log 2>&1
ncu -i /tmp/prof_cached.ncu-rep --page raw --csv > gpurun_out/prof_cached_raw.csv 2>/dev/null; ls -la gpurun_out/prof_cached_raw.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 80 --csv --log-file gpurun_out/launches_r14.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -12 gpurun_out/launches_r14.csv | cut -c1-220

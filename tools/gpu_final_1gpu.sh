#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --extras > gpurun_out/bench_extras.json 2> gpurun_out/bench_extras.err; echo "extras rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_default.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches','clocks')}); print('e2e',d['e2e']); print('roofline',d['roofline']); print('warm',d['warm_keycache']); print('nocache',d['no_keycache']); print('keyed',d['keyed']); print('kc',d['keycache']); print('cpu',d['cpu_baseline'])
r=json.load(open('gpurun_out/bench_reference.json')); print('reference arm',r['value'],r['cpu_baseline']['cores'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 30 -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"^k_(ed_verify_cached|kc_build|ed_hram)" -s 9 -c 3 -o /tmp/prof_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null; ls -la gpurun_out/prof_final_raw.csv gpurun_out/launches_final.csv

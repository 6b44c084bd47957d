#!/bin/bash
# round 2, run e: streaming SHA-256, constant-time signing, C harness, bundle test; bench with cfg4 in both signing modes
mkdir -p gpurun_out/r2e
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2e/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e/pytest_gpu.log
tail -25 gpurun_out/r2e/pytest_gpu.log
./tests/c_abi/harness; echo "harness rc=$?"
timeout 600 python bench.py --steps 50 > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('cfg4', d['cfg4']); print('cfg3', d['cfg3']); print('mix', d['issuer_mix'])
PY
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-secondary --extras > gpurun_out/r2e/bench_extras.json 2> gpurun_out/r2e/bench_extras.err; echo "extras rc=$?"
python -c "
import json; x=json.load(open('gpurun_out/r2e/bench_extras.json'))['extras']; print({k:(round(v.get('ms',0),3), v.get('signs_per_s')) for k,v in x.items() if k.startswith('sign')})"
AFC_SIGN_CT=0 timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-secondary --extras > gpurun_out/r2e/bench_extras_fast.json 2> gpurun_out/r2e/bench_extras_fast.err
python -c "
import json; x=json.load(open('gpurun_out/r2e/bench_extras_fast.json'))['extras']; print('fast', {k:(round(v.get('ms',0),3), v.get('signs_per_s')) for k,v in x.items() if k.startswith('sign')})"

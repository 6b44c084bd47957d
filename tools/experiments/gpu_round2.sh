#!/bin/bash
# Verify-kernel variants (I-cache / occupancy), parity re-check, extras + pipe probes.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for v in 0 1 2 3 4; do
  AFC_VERIFY_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_v$v.json"))
print("variant $v value %.3fM/s e2e %.3fM/s k_ed_verify %.3f ms hram %.3f ms clocks %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["kernel_avg_ms"], d["roofline"]["other_kernels_ms"]["k_ed_hram"], d["clocks"]))
PY
done
timeout 900 python bench.py --extras --steps 3 > gpurun_out/bench_extras.json 2> gpurun_out/bench_extras.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_extras.json')); print(json.dumps(d['extras'], indent=1)); print(d['value'], d['cpu_baseline'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ed_verify" -s 1 -c 1 -o gpurun_out/prof_verify_v1 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out

#!/bin/bash
mkdir -p gpurun_out
for lib in libafcrypto.so libafcrypto_ch32.so libafcrypto_ch64.so; do
  AFC_LIB=$PWD/agentfield_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$lib.json"))
print("$lib value %.2fM/s (%.3f ms) build %.3f ms verify %.3f ms" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["other_kernels_ms"]["k_kc_build"], d["roofline"]["kernel_avg_ms"]))
PY
done

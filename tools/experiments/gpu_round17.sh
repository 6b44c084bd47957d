#!/bin/bash
mkdir -p gpurun_out
for lib in libafcrypto.so libafcrypto_g8.so; do
  AFC_LIB=$PWD/agentfield_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$lib.json"))
print("$lib value %.2fM/s cached-kernel %.3f ms | no_keycache %.2fM/s (%.3f ms) | keyed %.2fM/s" % (d["value"]/1e6, d["roofline"]["kernel_avg_ms"], d["no_keycache"]["value"]/1e6, d["no_keycache"]["ms_per_step"], d["keyed"]["value"]/1e6))
PY
done

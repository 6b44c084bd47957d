#!/bin/bash
mkdir -p gpurun_out
for v in 0 2 3 4 5 1; do
  AFC_VERIFY_VARIANT=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_v$v.json"))
print("variant $v value %.3fM/s e2e %.3fM/s k_ed_verify %.3f ms hram %.3f ms" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["kernel_avg_ms"], d["roofline"]["other_kernels_ms"]["k_ed_hram"]))
PY
done

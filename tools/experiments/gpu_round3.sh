#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$name.json"))
print("$name value %.3fM/s e2e %.3fM/s k_ed_verify %.3f ms hram %.3f ms" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["roofline"]["kernel_avg_ms"], d["roofline"]["other_kernels_ms"]["k_ed_hram"]))
PY
}
run inline AFC_VERIFY_VARIANT=0
run call AFC_VERIFY_VARIANT=1
run kara_inline AFC_VERIFY_VARIANT=0 AFC_LIB=$PWD/agentfield_b200/libafcrypto_kara.so
run kara_call AFC_VERIFY_VARIANT=1 AFC_LIB=$PWD/agentfield_b200/libafcrypto_kara.so
timeout 900 python bench.py --extras --steps 3 --no-cpu-baseline > gpurun_out/bench_extras.json 2> gpurun_out/bench_extras.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_extras.json')); e=d['extras']
for k,v in e.items():
    if k!='microbench': print(k, v)
for k,v in e['microbench'].items(): print(k, v)"
AFC_LIB=$PWD/agentfield_b200/libafcrypto_kara.so python - <<'PY'
import agentfield_b200 as afb
c=afb.Context(0)
print("kara selftest", c.selftest(2000))
for name,w,it in (("fe_mul(kara)",0,4000),("fe_mul_schoolbook",7,4000),("fe_sq",1,4000)):
    print(name, c.microbench(w,it))
PY

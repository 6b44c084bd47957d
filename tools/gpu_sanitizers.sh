#!/bin/bash
# round 2: compute-sanitizer memcheck / racecheck / initcheck over the parity tests small enough to run under it — including the
# new code with shared memory and shuffles (k_kc_rows product tree, k_kc_chain4 quads, k_ed_sign_ct table staging, key-cache
# bookkeeping with eviction, streaming SHA-256, argument validation, go_cases, workflow VCs)
out=gpurun_out/r2san; mkdir -p $out
SEL="golden or ragged or empty or incremental or canonical or consistency or audit_proofs or codecs or expanded_key_cache_sign or keyed_verify or transparent_key_cache or constant_time or argument_validation or go_cases or workflow_vc"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "$SEL" > $out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $out/sanitizer_memcheck.log; tail -4 $out/sanitizer_memcheck.log | cut -c1-200
for tool in racecheck initcheck; do
timeout 1200 compute-sanitizer --tool $tool --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
  -k "golden or ragged or empty or incremental or consistency or audit_proofs or codecs or keyed_verify or transparent_key_cache or constant_time or argument_validation" > $out/sanitizer_$tool.log 2>&1
echo "$tool rc=$?" >> $out/sanitizer_$tool.log; tail -4 $out/sanitizer_$tool.log | cut -c1-200
done
# the kernels that are not the default: static split, four lanes per credential (shuffles + shared-memory slots)
for knob in AFC_VERIFY_DYNAMIC=0 AFC_VERIFY_QUAD=2; do
for tool in memcheck racecheck; do
env $knob timeout 900 compute-sanitizer --tool $tool --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
  -k "edge_set or keyed_verify or transparent_key_cache" > $out/sanitizer_${tool}_$knob.log 2>&1
echo "$tool $knob rc=$?" >> $out/sanitizer_${tool}_$knob.log; tail -3 $out/sanitizer_${tool}_$knob.log | cut -c1-200
done; done
grep -c "ERROR SUMMARY" $out/*.log; grep "ERROR SUMMARY\|RACECHECK SUMMARY" $out/*.log | head

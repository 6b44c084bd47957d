#!/bin/bash
# memory-safety evidence: compute-sanitizer memcheck over the parity tests that are small enough to run under it
mkdir -p gpurun_out
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
  -k "golden or ragged or empty or incremental or canonical or consistency or audit_proofs or codecs or expanded_key_cache_sign or keyed_verify or transparent_key_cache" \
  > gpurun_out/sanitizer_memcheck.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitizer_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitizer_memcheck.log; tail -6 gpurun_out/sanitizer_memcheck.log

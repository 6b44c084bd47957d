#!/bin/bash
mkdir -p gpurun_out
for tool in racecheck initcheck; do
timeout 500 compute-sanitizer --tool $tool --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
  -k "golden or ragged or empty or incremental or canonical or consistency or audit_proofs or codecs or expanded_key_cache_sign or keyed_verify" \
  > gpurun_out/sanitizer_$tool.log 2>&1
echo "$tool rc=$?" >> gpurun_out/sanitizer_$tool.log
tail -4 gpurun_out/sanitizer_$tool.log | cut -c1-200
done

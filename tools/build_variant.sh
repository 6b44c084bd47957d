#!/bin/bash
# build_variant.sh NAME "-DFLAG=... ..."  ->  agentfield_b200/variants/libafcrypto_NAME.so  (experiment builds; load with AFC_LIB=...)
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
obj=$root/build/variants/$name; mkdir -p $obj $root/agentfield_b200/variants
cd $root/agentfield_b200/csrc
for f in k_hash k_ed25519 afcrypto afc_ingest; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags -c $f.cu -o $obj/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $root/agentfield_b200/variants/libafcrypto_$name.so $obj/k_hash.o $obj/k_ed25519.o $obj/afcrypto.o $obj/afc_ingest.o -ldl -lpthread
echo built $name

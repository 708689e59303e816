#!/bin/bash
# build a tuning variant of libsnpgpu: profiles/variant.sh <name> <nvcc -D flags...>  ->  gpurun_out/../variants/libsnpgpu_<name>.so
set -e
name=$1; shift
cd "$(dirname "$0")/../snp_pipeline_b200/csrc"
mkdir -p ../../variants/obj_$name
for f in api k1_pileup k2_merge k3_sites k4_distance k5_vcf k6_metrics k7_regions synth; do
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-Wno-deprecated-declarations -Wno-deprecated-declarations "$@" -c $f.cu -o ../../variants/obj_$name/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libsnpgpu_$name.so ../../variants/obj_$name/*.o -lcudart
rm -rf ../../variants/obj_$name

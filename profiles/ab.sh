#!/bin/bash
# A/B on the box: for each variant name ("base" = the in-tree library) run the GPU parity tests and K1 alone in both modes
tag=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset SNPGPU_LIB; else export SNPGPU_LIB=$PWD/variants/libsnpgpu_$v.so; fi
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_${v}_pytest.log 2>&1; echo "$v: $(tail -1 gpurun_out/${tag}_${v}_pytest.log)"
  python profiles/run_k1.py all 8 > gpurun_out/${tag}_${v}_all.log 2>&1; echo "$v all:   $(tail -1 gpurun_out/${tag}_${v}_all.log)"
  python profiles/run_k1.py sites 8 > gpurun_out/${tag}_${v}_sites.log 2>&1; echo "$v sites: $(tail -1 gpurun_out/${tag}_${v}_sites.log)"
done

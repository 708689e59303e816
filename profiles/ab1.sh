#!/bin/bash
# timing-only A/B, one round: profiles/ab1.sh <variant names...> ("base" = in-tree)
for v in "$@"; do
  if [ "$v" = base ]; then unset SNPGPU_LIB; else export SNPGPU_LIB=$PWD/variants/libsnpgpu_$v.so; fi
  echo "$v all:   $(python profiles/run_k1.py all 6 2>&1 | tail -1 | cut -c1-150)"
  echo "$v sites: $(python profiles/run_k1.py sites 6 2>&1 | tail -1 | cut -c1-150)"
done

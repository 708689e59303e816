#!/bin/bash
# K1 (all-positions mode) against the share of lines that carry an indel token: profiles/indel_sweep.sh <variant names...> ("base" = in-tree)
for v in "$@"; do
  if [ "$v" = base ]; then unset SNPGPU_LIB; else export SNPGPU_LIB=$PWD/variants/libsnpgpu_$v.so; fi
  for r in 0 0.003 0.01 0.03 0.1; do
    echo "$v indel_line_rate $r: $(INDEL_RATE=$r python profiles/run_k1.py all 6 2>&1 | tail -1 | cut -c1-170)"
  done
done

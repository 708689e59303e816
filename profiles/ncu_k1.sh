#!/bin/bash
# one full ncu capture of the pileup kernel and its follow-up kernel (tag = $1, mode = $2: all | sites): raw metrics +
# source-level pages come back in gpurun_out/
tag=$1; mode=${2:-all}
BATCH=2 ncu --set full --clock-control none --import-source on -k regex:'k1_pileup_kernel|k1_rest_kernel' -s 2 -c 2 -f -o gpurun_out/k1_${tag} \
    python profiles/run_k1.py $mode 3 > gpurun_out/${tag}_ncu.log 2>&1; tail -1 gpurun_out/${tag}_ncu.log

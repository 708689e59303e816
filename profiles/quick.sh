#!/bin/bash
# quick on-box check: GPU parity tests, then K1 alone in both modes (tag = $1)
tag=$1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
python profiles/run_k1.py all 6 > gpurun_out/${tag}_k1_all.log 2>&1; tail -1 gpurun_out/${tag}_k1_all.log
python profiles/run_k1.py sites 6 > gpurun_out/${tag}_k1_sites.log 2>&1; tail -1 gpurun_out/${tag}_k1_sites.log
BATCH=1 python profiles/run_k1.py all 10 > gpurun_out/${tag}_k1_all_b1.log 2>&1; tail -1 gpurun_out/${tag}_k1_all_b1.log

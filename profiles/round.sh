#!/bin/bash
# Everything a round's evidence needs, on the GPU box: profiles/round.sh <tag>
#   GPU tests, smoke, the bench (both arms), the ncu launch list of a short bench run, one full ncu capture of K1.
tag=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; cut -c1-300 gpurun_out/${tag}_bench_ref.json
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-1200 gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --samples 8 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${tag}_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1_pileup -s 2 -c 1 -f -o gpurun_out/k1_${tag} \
    python profiles/run_k1.py all 4 > gpurun_out/${tag}_ncu.log 2>&1; tail -1 gpurun_out/${tag}_ncu.log

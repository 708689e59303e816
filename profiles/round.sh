#!/bin/bash
# Everything a round's evidence needs, on the GPU box: profiles/round.sh <tag>
#   GPU tests, smoke, the bench (both arms), the ncu launch list of a short bench run, full ncu captures of K1 (both modes)
#   and K4, the indel-rate sensitivity of K1.
tag=$1
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
python bench.py --impl reference > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; cut -c1-300 gpurun_out/${tag}_bench_ref.json
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cut -c1-600 gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --samples 16 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${tag}_launch_bench.log 2>&1
bash profiles/ncu_k1.sh ${tag} all
bash profiles/ncu_k1.sh ${tag}s sites
if [ -z "$SKIP_K4" ]; then
ncu --set full --clock-control none --import-source on -k regex:k4_ -s 2 -c 2 -f -o gpurun_out/k4_${tag} \
    python profiles/run_k4.py 5000 200000 625 2 > gpurun_out/${tag}_k4_ncu.log 2>&1; tail -2 gpurun_out/${tag}_k4_ncu.log
fi
python profiles/step_timeline.py > gpurun_out/${tag}_step_timeline.txt 2>&1; grep -E "^span" gpurun_out/${tag}_step_timeline.txt
for r in 0 0.003 0.01 0.03 0.1; do echo "indel_line_rate $r: $(INDEL_RATE=$r python profiles/run_k1.py all 5 2>&1 | tail -1 | cut -c1-110)"; done | tee gpurun_out/${tag}_indel.txt
echo "sites, 200k-site union: $(EXTRA_SITES=200000 python profiles/run_k1.py sites 5 2>&1 | tail -1 | cut -c1-110)" | tee -a gpurun_out/${tag}_indel.txt

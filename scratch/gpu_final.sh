#!/bin/bash
# final single-GPU evidence: default bench (both arms), launch list, ncu --set full of the kernels that changed
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/ncu_l.log 2>&1
python scratch/agg_launches.py gpurun_out/launches.csv 40 > gpurun_out/launches_summary.txt; head -12 gpurun_out/launches_summary.txt
for K in k_rho_tiles k_rho_fold k_nb_fill k_pair; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:$K --launch-skip 1 --launch-count 1 -o gpurun_out/fin_$K -f python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_$K.log 2>&1
  echo "$K ncu rc=$?"
done

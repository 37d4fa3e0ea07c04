#!/bin/bash
# Run on the GPU box (gpurun): ncu captures of the two sweep kernels + the launch list of a short bench run.
# Outputs under gpurun_out/; tools/ncu_summary.py / ncu_source_top.py turn them into the text files kept in profiles/.
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:^sweep_l2 -c 1 -o gpurun_out/prof_l2_final \
    python tools/profile_step.py surf 38 8000 1 > gpurun_out/ncu_l2_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^sweep_hamming -c 1 -o gpurun_out/prof_ham_final \
    python tools/profile_step.py orb 60 4000 1 > gpurun_out/ncu_ham_final.log 2>&1
ncu --set full --clock-control none -k regex:^finalize -c 2 -o gpurun_out/prof_finalize_final \
    python tools/profile_step.py surf 38 8000 1 > gpurun_out/ncu_fin_final.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(sweep_|finalize|pack_)' -c 80 --csv \
    --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --cpu-budget-s 0 > gpurun_out/bench_under_ncu.log 2>&1
tail -5 gpurun_out/launches_final.csv | cut -c1-200

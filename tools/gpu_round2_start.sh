#!/bin/bash
# First GPU call of round 2: everything round 1 could not re-measure after the ORB "Z" encoding became the default.
# Every command under its own timeout; everything lands in gpurun_out/ as it finishes.
# Afterwards, here: python tools/profiles_from_gpurun.py r2   (gpurun_out/ -> profiles/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 400 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
( time timeout 300 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json
ORB_Z_PROBE_MODES=0,1 ORB_Z_PROBE_DRAIN=1 timeout 120 python tools/orb_z_probe.py > gpurun_out/orb_z_probe.txt 2>&1; cat gpurun_out/orb_z_probe.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:^sweep_l2_tc -c 1 -f -o gpurun_out/prof_ham_z \
    python tools/profile_step.py orb 60 4000 1 > gpurun_out/ncu_ham_z.log 2>&1; tail -2 gpurun_out/ncu_ham_z.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^(sweep_|finalize|pack_)' -c 120 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-budget-s 0 --no-alt-engine > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/launches.csv | cut -c1-200
timeout 200 python tools/surf_bf_probe.py > gpurun_out/surf_bf_probe.txt 2>&1; cat gpurun_out/surf_bf_probe.txt

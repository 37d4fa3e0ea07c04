#!/bin/bash
# Last GPU call of round 1 (a few minutes of box time): ORB tensor-core probes, ncu capture of the ORB tensor-core sweep,
# the default bench line, launch list + DRAM traffic of a short bench under ncu, then the GPU tests with what is left.
# Every command under its own timeout; everything lands in gpurun_out/ as it finishes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 150 python tools/orb_tc_probe.py > gpurun_out/orb_tc_probe.txt 2>&1; tail -9 gpurun_out/orb_tc_probe.txt
timeout 150 ncu --set full --clock-control none --import-source on -k regex:^sweep_l2_tc -c 1 -f -o gpurun_out/prof_ham_tc \
    python tools/profile_step.py orb 60 4000 1 > gpurun_out/ncu_ham_tc.log 2>&1; tail -2 gpurun_out/ncu_ham_tc.log
( time timeout 240 python bench.py --cpu-budget-s 8 ) > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^(sweep_|finalize|pack_)' -c 120 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-budget-s 0 --no-alt-engine > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/launches.csv | cut -c1-200
( ESFM_TC_QT_ORB=2 timeout 200 python -m pytest tests -m gpu -x -q -k "tc and (hamming or orb or kat or fountain or synth or edge or mutual or persist)" ) > gpurun_out/pytest_gpu_orb_qt2.log 2>&1
tail -3 gpurun_out/pytest_gpu_orb_qt2.log
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log

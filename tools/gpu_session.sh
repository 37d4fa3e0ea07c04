#!/bin/bash
# Round validation call: GPU parity tests, the default bench line, launch list + DRAM traffic of a short bench under ncu,
# full ncu capture of the TC sweep.  Every command under its own timeout; everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 90 python tools/profile_step.py surf 12 2000 1 tc > gpurun_out/smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/smoke.txt; exit 1; }
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( time BENCH_E2E_DEBUG=1 timeout 300 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json; grep "e2e step" gpurun_out/bench.err | tail -4
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke_entry.txt 2>&1; tail -2 gpurun_out/smoke_entry.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^(sweep_|finalize|pack_)' -c 120 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-budget-s 0 > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches.csv | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^sweep_l2_tc -c 1 -f -o gpurun_out/prof_l2_tc \
    python tools/profile_step.py surf 38 8000 1 tc > gpurun_out/ncu_l2_tc.log 2>&1

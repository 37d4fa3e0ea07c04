#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line, TC-sweep pipeline probes, ncu captures (full set of the TC
# sweep + launch list of a short bench).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-600 gpurun_out/bench.json
for d in 0 1 2 4 3 6; do
  ESFM_TC_DEBUG=$d timeout 120 python tools/profile_step.py surf 38 8000 3 tc 2>&1 | tail -1 | sed "s/^/debug=$d /" >> gpurun_out/tc_probes.txt
done
cat gpurun_out/tc_probes.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:^sweep_l2_tc -c 1 -f -o gpurun_out/prof_l2_tc \
    python tools/profile_step.py surf 38 8000 1 tc > gpurun_out/ncu_l2_tc.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(sweep_|finalize|pack_)' -c 120 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-budget-s 0 > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches.csv | cut -c1-200

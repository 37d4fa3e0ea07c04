#!/bin/bash
# The one GPU-session script (run on the box through gpurun):  bash tools/gpu.sh <stage> [<stage> ...]
# Every stage runs under its own timeout and leaves its output under gpurun_out/; afterwards, here:
#   python tools/profiles_from_gpurun.py r2      (gpurun_out/ -> profiles/)
# Stages:
#   info        GPU name / clocks / power limit
#   smoke       __graft_entry__.smoke()  (aborts the session on failure: nothing else is worth GPU time then)
#   tests       pytest -m gpu            (TESTS_K="expr" restricts with -k)
#   bench       python bench.py          (BENCH_ARGS="..." appended)
#   launches    short bench under ncu: per-launch device time + DRAM bytes (launches.csv)
#   ncu_surf    ncu --set full of the SURF tensor-core sweep (ENGINE=tc16|tc) at 2346 pairs of 8000 x 8000 (units_per_pair == 1)
#   ncu_orb     ncu --set full of the ORB tensor-core sweep (ENGINE=tc16|tc) at 9453 pairs of 4000 x 4000
#   zprobe      tools/orb_z_probe.py with the drain-only probe
#   probes      pipeline probes ($ESFM_TC_DEBUG) of both tensor-core sweeps
#   multi       N-device tests + N-rank bench (N = $N, default 2): use with gpurun --gpus N
#   fulljob     tools/full_job.py on $N devices (KINDS="surf orb")
#   sanitizer   compute-sanitizer memcheck + racecheck on the smoke shapes
#   orb         ORB extraction: tools/orb_bench.py (vs cv2 on the host cores) + ncu launch list of one 1080p frame
mkdir -p gpurun_out
N=${N:-2}
for stage in "$@"; do
case $stage in
info)
  nvidia-smi --query-gpu=index,name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1; cat gpurun_out/gpu.txt
  nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt ;;
smoke)
  timeout 180 python __graft_entry__.py smoke > gpurun_out/smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -15 gpurun_out/smoke.txt; exit 1; }
  tail -2 gpurun_out/smoke.txt | cut -c1-200 ;;
tests)
  ( time timeout ${TESTS_TIMEOUT:-600} python -m pytest tests -m gpu ${TESTS_FLAGS--x} -q --durations=10 ${TESTS_K:+-k "$TESTS_K"} ) > gpurun_out/pytest_gpu.log 2>&1
  tail -25 gpurun_out/pytest_gpu.log ;;
bench)
  ( time BENCH_E2E_DEBUG=1 timeout 420 python bench.py $BENCH_ARGS ) > gpurun_out/bench.json 2> gpurun_out/bench.err
  cut -c1-400 gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
launches)
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:^(sweep_|finalize|pack_)" -c 120 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-budget-s 0 --no-alt-engine --no-parity > gpurun_out/bench_under_ncu.log 2>&1
  tail -2 gpurun_out/launches.csv | cut -c1-200 ;;
ncu_surf)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^sweep_ -c 1 -f -o gpurun_out/prof_surf_${ENGINE:-tc16} \
      python tools/profile_step.py surf 69 8000 1 ${ENGINE:-tc16} > gpurun_out/ncu_surf.log 2>&1; tail -2 gpurun_out/ncu_surf.log | cut -c1-200 ;;
ncu_orb)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^sweep_ -c 1 -f -o gpurun_out/prof_orb_${ENGINE:-tc16} \
      python tools/profile_step.py orb 138 4000 1 ${ENGINE:-tc16} > gpurun_out/ncu_orb.log 2>&1; tail -2 gpurun_out/ncu_orb.log | cut -c1-200 ;;
zprobe)
  ORB_Z_PROBE_MODES=${ORB_Z_PROBE_MODES:-1} ORB_Z_PROBE_DRAIN=1 timeout 150 python tools/orb_z_probe.py > gpurun_out/orb_z_probe.txt 2>&1; cat gpurun_out/orb_z_probe.txt ;;
probes)
  rm -f gpurun_out/tc_probes.txt
  for d in ${PROBE_FLAGS:-0 1 8 16 24}; do
    ESFM_TC_DEBUG=$d timeout 60 python tools/profile_step.py surf 38 8000 3 ${ENGINE:-tc16} 2>&1 | tail -1 | sed "s/^/surf debug=$d /" >> gpurun_out/tc_probes.txt
    ESFM_TC_DEBUG=$d timeout 60 python tools/profile_step.py orb 60 4000 3 ${ENGINE:-tc16} 2>&1 | tail -1 | sed "s/^/orb  debug=$d /" >> gpurun_out/tc_probes.txt
  done
  cat gpurun_out/tc_probes.txt ;;
multi)
  ( timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi.log 2>&1; tail -4 gpurun_out/pytest_multi.log
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 6 --warmup 3 --no-secondary \
     > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  cut -c1-600 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err ;;
fulljob)
  for k in ${KINDS:-surf orb}; do
    timeout ${FULLJOB_TIMEOUT:-600} python tools/full_job.py $k --gpus $N $FULLJOB_ARGS > gpurun_out/full_job_${k}_${N}gpu.json 2> gpurun_out/full_job_${k}_${N}gpu.err
    cut -c1-700 gpurun_out/full_job_${k}_${N}gpu.json; tail -3 gpurun_out/full_job_${k}_${N}gpu.err
  done ;;
sanitizer)
  for tool in memcheck racecheck; do
    timeout ${SAN_TIMEOUT:-240} compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer_$tool.log 2>&1
    echo "compute-sanitizer $tool exit $?"; tail -4 gpurun_out/sanitizer_$tool.log | cut -c1-200
  done ;;
orb)
  timeout 200 python tools/orb_bench.py --frames ${ORB_FRAMES:-12} 2>&1 | tail -4 | cut -c1-700
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/orb_launches.csv python tools/orb_profile.py 2>&1 | tail -1 ;;
*)
  if [ -f "$stage" ]; then timeout ${SCRIPT_TIMEOUT:-300} python "$stage" > "gpurun_out/$(basename "$stage" .py).txt" 2>&1; tail -30 "gpurun_out/$(basename "$stage" .py).txt"
  else echo "unknown stage $stage"; fi ;;
esac
done

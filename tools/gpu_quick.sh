#!/bin/bash
# quick iteration call: smoke (aborts the call on failure), GPU parity tests, pipeline probes of the TC sweep (both geometries).
# Every command runs under its own short timeout: a hung kernel must not eat the GPU budget.
mkdir -p gpurun_out
rm -f gpurun_out/tc_probes.txt
for qt in 1 2; do
  ESFM_TC_QT=$qt timeout 90 python tools/profile_step.py surf 12 2000 1 tc > gpurun_out/smoke_qt$qt.txt 2>&1 || { echo "SMOKE qt=$qt FAILED"; tail -5 gpurun_out/smoke_qt$qt.txt; exit 1; }
  tail -1 gpurun_out/smoke_qt$qt.txt
done
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
( ESFM_TC_QT=2 timeout 200 python -m pytest tests -m gpu -x -q -k "tc" ) > gpurun_out/pytest_gpu_qt2.log 2>&1
tail -4 gpurun_out/pytest_gpu_qt2.log
for qt in 1 2; do for d in 0 1 5; do
  ESFM_TC_QT=$qt ESFM_TC_DEBUG=$d timeout 60 python tools/profile_step.py surf 38 8000 3 tc 2>&1 | tail -1 | sed "s/^/qt=$qt debug=$d /" >> gpurun_out/tc_probes.txt
done; done
cat gpurun_out/tc_probes.txt
for qt in 1 2; do
ESFM_TC_QT=$qt timeout 120 python bench.py --no-secondary --no-alt-engine --cpu-budget-s 0 > gpurun_out/bench_surf_qt$qt.json 2> gpurun_out/bench_surf_qt$qt.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_surf_qt$qt.json'))
    print('qt=$qt value %.4g e2e %.4g kernel_ms %.2f clocks %s frac_exec %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['clocks'], d['roofline'].get('frac_executed',0)))
except Exception as e: print('bench qt=$qt failed', e)
PY
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^sweep_l2_tc -c 1 -f -o gpurun_out/prof_l2_tc \
    python tools/profile_step.py surf 38 8000 1 tc > gpurun_out/ncu_l2_tc.log 2>&1

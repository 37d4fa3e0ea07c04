#!/bin/bash
# iteration call: smoke, dense-row threshold sweep ($ESFM_TC_DEBUG bits 8-15), full GPU tests, default bench
mkdir -p gpurun_out
rm -f gpurun_out/tc_probes.txt
timeout 90 python tools/profile_step.py surf 12 2000 1 tc > gpurun_out/smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/smoke.txt; exit 1; }
tail -1 gpurun_out/smoke.txt
for dl in 4 8 16 40; do
  ESFM_TC_DEBUG=$((dl * 256)) timeout 60 python tools/profile_step.py surf 38 8000 3 tc 2>&1 | tail -1 | sed "s/^/dense_lanes=$dl /" >> gpurun_out/tc_probes.txt
done
cat gpurun_out/tc_probes.txt
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( time BENCH_E2E_DEBUG=1 timeout 300 python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value %.4g e2e %.4g kernel_ms %.2f clocks %s frac_exec %.3f orb %.4g' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['clocks'], d['roofline'].get('frac_executed',0), d['secondary']['value']))
PY
grep "e2e step" gpurun_out/bench.err | sed -n 5,6p

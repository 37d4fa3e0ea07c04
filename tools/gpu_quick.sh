#!/bin/bash
# quick iteration call: tensor-core probe, GPU parity tests, pipeline probes of the TC sweep
mkdir -p gpurun_out
timeout 120 easysfm_b200/bin/tc_probe > gpurun_out/tc_probe.txt 2>&1; tail -12 gpurun_out/tc_probe.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
rm -f gpurun_out/tc_probes.txt
for d in 0 1 4 5; do
  ESFM_TC_DEBUG=$d timeout 120 python tools/profile_step.py surf 38 8000 3 tc 2>&1 | tail -1 | sed "s/^/debug=$d /" >> gpurun_out/tc_probes.txt
done
cat gpurun_out/tc_probes.txt
timeout 300 python bench.py --no-secondary --no-alt-engine --cpu-budget-s 0 > gpurun_out/bench_surf.json 2> gpurun_out/bench_surf.err
cut -c1-200 gpurun_out/bench_surf.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_surf.json'))
print('value %.4g e2e %.4g kernel_ms %.2f clocks %s frac_exec %.3f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['clocks'], d['roofline'].get('frac_executed',0)))
PY

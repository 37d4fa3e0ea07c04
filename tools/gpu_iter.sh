#!/bin/bash
# shortest iteration call: smoke (aborts on failure), pipeline probes, one ncu capture.  Every command under its own timeout.
mkdir -p gpurun_out
rm -f gpurun_out/tc_probes.txt
ESFM_TC_QT=${QT:-1} timeout 90 python tools/profile_step.py surf 12 2000 1 tc > gpurun_out/smoke.txt 2>&1 || { echo "SMOKE FAILED"; tail -5 gpurun_out/smoke.txt; exit 1; }
tail -1 gpurun_out/smoke.txt
for qt in ${QTS:-1 2}; do for d in ${DBG:-0 1}; do
  ESFM_TC_QT=$qt ESFM_TC_DEBUG=$d timeout 60 python tools/profile_step.py surf 38 8000 3 tc 2>&1 | tail -1 | sed "s/^/qt=$qt debug=$d /" >> gpurun_out/tc_probes.txt
done; done
cat gpurun_out/tc_probes.txt
if [ "${NCU:-1}" = "1" ]; then
ESFM_TC_QT=${QT:-1} timeout 200 ncu --set full --clock-control none --import-source on -k regex:^sweep_l2_tc -c 1 -f -o gpurun_out/prof_l2_tc \
    python tools/profile_step.py surf 38 8000 1 tc > gpurun_out/ncu_l2_tc.log 2>&1
fi

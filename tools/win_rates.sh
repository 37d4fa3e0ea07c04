#!/bin/bash
# rates of the tc16 sweeps next to the fp32-accumulator ones (703 SURF pairs of 8000 x 8000 / 1770 ORB pairs of 4000 x 4000), with pipeline probes
for d in ${PROBE_FLAGS:-0}; do
  for e in tc tc16; do
    ESFM_TC_DEBUG=$d timeout 90 python tools/profile_step.py surf 38 8000 3 $e 2>&1 | tail -1 | sed "s/^/debug=$d /"
    ESFM_TC_DEBUG=$d timeout 90 python tools/profile_step.py orb 60 4000 3 $e 2>&1 | tail -1 | sed "s/^/debug=$d /"
  done
done

mkdir -p gpurun_out
( time timeout 60 python -m pytest tests -m gpu -x -q -k "not config2" ) > gpurun_out/pytest_gpu_z.log 2>&1; tail -4 gpurun_out/pytest_gpu_z.log
timeout 25 python __graft_entry__.py smoke > gpurun_out/smoke_z.txt 2>&1; tail -2 gpurun_out/smoke_z.txt
ORB_Z_PROBE_MODES=1,2 timeout 40 python tools/orb_z_probe.py > gpurun_out/orb_z_probe2.txt 2>&1; cat gpurun_out/orb_z_probe2.txt

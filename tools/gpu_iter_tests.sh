bash tools/gpu_iter.sh || exit 1
( timeout 200 python -m pytest tests -m gpu -x -q -k "tc" ) > gpurun_out/pytest_gpu_tc.log 2>&1; tail -3 gpurun_out/pytest_gpu_tc.log

mkdir -p gpurun_out
timeout 150 python tools/ham_tc_check.py > gpurun_out/ham_tc_check.txt 2>&1; tail -12 gpurun_out/ham_tc_check.txt

#!/bin/bash
# N-GPU call (gpurun --gpus N): byte-identical sharded results, then the bench line at N ranks.
N=${N:-2}
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi.log 2>&1; tail -2 gpurun_out/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 6 --warmup 3 --no-secondary \
   > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cut -c1-400 gpurun_out/bench_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 4 --warmup 3 --no-secondary --kind orb \
   > gpurun_out/bench_orb_n$N.json 2> gpurun_out/bench_orb_n$N.err
cut -c1-300 gpurun_out/bench_orb_n$N.json

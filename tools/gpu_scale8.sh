#!/bin/bash
# Run ON an N-GPU box (gpurun --gpus N): BASELINE.json configs[3] (1024^2, 304 frames per clip, batches of 19; weak `value` + the
# `strong` block = ONE clip sharded over the N ranks) and configs[4] (training step, NCCL all-reduce inside the CUDA graph).
# The default config at N = 1/2/4/8 is the driver's own SCALE run.
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 300 $RUN --master-port 29512 bench.py --gpus $N --size 1024 --frames 304 --batch 19 --steps 2 --warmup 2 --no-png --no-lib-baseline --no-cpu-baseline > gpurun_out/cfg4_n${N}.json 2> gpurun_out/cfg4_n${N}.err)
head -c 1200 gpurun_out/cfg4_n${N}.json; echo; tail -2 gpurun_out/cfg4_n${N}.err
(timeout 200 $RUN --master-port 29513 bench.py --gpus $N --config train --steps 5 --warmup 2 > gpurun_out/train_n${N}.json 2> gpurun_out/train_n${N}.err)
head -c 600 gpurun_out/train_n${N}.json; echo; tail -2 gpurun_out/train_n${N}.err

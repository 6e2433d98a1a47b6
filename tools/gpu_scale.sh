#!/bin/bash
# Run ON an N-GPU box (gpurun --gpus N): the bench contract at N ranks (weak `value` + the `strong` block: one 300-frame clip split
# over the ranks) and BASELINE.json configs[3] (1024^2, 304 frames in contiguous shards, batches of 19), plus the training step.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 900 $RUN --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/scale_n${N}.json 2> gpurun_out/scale_n${N}.err)
head -c 1500 gpurun_out/scale_n${N}.json; echo; tail -3 gpurun_out/scale_n${N}.err
(timeout 900 $RUN --master-port 29512 bench.py --gpus $N --size 1024 --frames 304 --batch 19 --steps 2 --warmup 2 --no-png --no-lib-baseline > gpurun_out/cfg4_n${N}.json 2> gpurun_out/cfg4_n${N}.err)
head -c 1200 gpurun_out/cfg4_n${N}.json; echo; tail -3 gpurun_out/cfg4_n${N}.err
(timeout 600 $RUN --master-port 29513 bench.py --gpus $N --config train --steps 3 --warmup 2 > gpurun_out/train_n${N}.json 2> gpurun_out/train_n${N}.err)
head -c 800 gpurun_out/train_n${N}.json; echo; tail -3 gpurun_out/train_n${N}.err

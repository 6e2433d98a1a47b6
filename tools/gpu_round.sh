#!/bin/bash
# One GPU-box session: tests, bench (ours + reference arm + training config), ncu evidence.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -200 > gpurun_out/pytest_gpu.log)
tail -15 gpurun_out/pytest_gpu.log
(timeout 600 python bench.py --steps 3 --warmup 3 --dump-layers gpurun_out/layers.json > gpurun_out/bench1.json 2> gpurun_out/bench1.err)
head -c 2500 gpurun_out/bench1.json; echo
(timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err)
(timeout 400 python bench.py --config train --steps 3 --warmup 2 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err)
head -c 600 gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
if [ "${1:-}" = "profile" ]; then bash tools/gpu_profile.sh r02 quick > gpurun_out/profile.log 2>&1; tail -5 gpurun_out/profile.log; fi

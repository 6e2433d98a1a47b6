#!/bin/bash
# Run ON THE GPU BOX: ncu evidence for the training-step GEMM kernels (csrc/train.cu) in the second eager step of bench.py --config train.
#   usage: tools/gpu_profile_train.sh [tag]  -> gpurun_out/<tag>_train_gemm_ncu.csv (one row per profiled launch)
TAG=${1:-r02}
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none -k regex:train_gemm --launch-skip 520 -c 240 -o /tmp/ncu/${TAG}_train \
    python bench.py --config train --no-graph --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_train.log 2>&1
echo "ncu exit $?"
python tools/ncu_extract.py /tmp/ncu/${TAG}_train.ncu-rep gpurun_out/${TAG}_train_gemm_ncu.csv
wc -l gpurun_out/${TAG}_train_gemm_ncu.csv

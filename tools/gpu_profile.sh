#!/bin/bash
# Run ON THE GPU BOX (via gpurun): per-launch device times of two steady-state 20-frame batches, and one full ncu
# capture of the dominant conv kernel.  Outputs land in gpurun_out/ (summaries are copied to profiles/ by hand).
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 1 --frames 20 --batch 20 --no-graph --no-cpu-baseline"
# iper kernels only; skip source setup (35 launches) + engine warm-up batch (71) => two full batches follow
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:iper:: -s 106 -c 142 --csv \
    --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
# dominant kernel: 3x3 256->256 residual conv (BN=256, split fp16): skip the 12+2 source-side launches of that variant
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 60 -c 6 \
    -o gpurun_out/conv_full $BENCH > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 1 -c 1 \
    -o gpurun_out/raster_full $BENCH > gpurun_out/ncu_raster.log 2>&1
echo "raster capture exit $?"
ls -la gpurun_out

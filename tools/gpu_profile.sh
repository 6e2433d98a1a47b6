#!/bin/bash
# Run ON THE GPU BOX (via gpurun): per-launch device times of steady-state 20-frame batches and full ncu captures of the
# kernels the roofline claims rest on.  Reports are trimmed to CSV on the box (tools/ncu_extract.py) because gpurun only
# returns 64 MiB; one small .ncu-rep (a residual-block conv, with source) is kept for source-level reading.
set -u
mkdir -p gpurun_out /tmp/ncu
BENCH="python bench.py --steps 1 --warmup 1 --frames 20 --batch 20 --no-graph --no-cpu-baseline"
# iper kernels only; skip source setup (35 launches) + engine warm-up batch (63) => full batches follow
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:iper:: \
    -s 98 -c 130 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
# conv stack of one batch: skip 14 forward_src convs + 9 source-map projections + the 38 of the warm-up batch
timeout 1200 ncu --set full --clock-control none -k "regex:conv_gemm|conv_halo" -s 61 -c 38 -o /tmp/ncu/conv_all $BENCH \
    > gpurun_out/ncu_conv.log 2>&1
echo "conv capture exit $?"
python tools/ncu_extract.py /tmp/ncu/conv_all.ncu-rep gpurun_out/conv_gemm_ncu.csv
if [ "${1:-}" = "quick" ]; then ls -la gpurun_out; exit 0; fi     # quick: launch list + conv capture only
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_gemm|conv_halo" -s 69 -c 1 \
    -o gpurun_out/conv_res0b $BENCH > gpurun_out/ncu_conv1.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:raster_kernel|raster_setup_kernel|warp_attention_kernel|conv_stem_kernel|flow_resize_kernel|pred_to_u8_kernel" \
    -s 15 -c 17 -o /tmp/ncu/hbm_all $BENCH > gpurun_out/ncu_hbm.log 2>&1
echo "hbm-kernel capture exit $?"
python tools/ncu_extract.py /tmp/ncu/hbm_all.ncu-rep gpurun_out/hbm_kernels_ncu.csv
ls -la gpurun_out

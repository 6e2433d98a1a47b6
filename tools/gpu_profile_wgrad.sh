#!/bin/bash
# Run ON THE GPU BOX: a light ncu pass (11 metrics, not --set full) over the MN-major weight-gradient launches of the second eager
# training step -> gpurun_out/<tag>_train_wgrad_ncu.csv
TAG=${1:-r02}
mkdir -p gpurun_out /tmp/ncu
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
M=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__cycles_elapsed.max
timeout 240 ncu --metrics $M --clock-control none --kernel-name-base demangled -k "regex:train_gemm_kernel<.int.[0-9]+,..int.1>" --launch-skip 140 -c 90 -o /tmp/ncu/${TAG}_wgrad \
    python bench.py --config train --no-graph --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu_wgrad.log 2>&1
echo "ncu exit $?"
python tools/ncu_extract.py /tmp/ncu/${TAG}_wgrad.ncu-rep gpurun_out/${TAG}_train_wgrad_ncu.csv

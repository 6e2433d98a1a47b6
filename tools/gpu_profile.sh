#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu evidence for one steady-state 20-frame batch of the bench workload (512^2, ns=2, fp16x2).
# bench.py --profile-batch brackets exactly one eager batch with cudaProfilerStart/Stop, so `--profile-from-start off` captures
# that batch and nothing else (no source setup, no warm-up), whatever the launch count of the build is.
# Reports are trimmed to CSV on the box (tools/ncu_extract.py): gpurun returns at most 64 MiB.
#   usage: tools/gpu_profile.sh [tag]      -> gpurun_out/<tag>_launches.csv, <tag>_all_kernels_ncu.csv, <tag>_conv_res.ncu-rep
set -u
TAG=${1:-r02}
mkdir -p gpurun_out /tmp/ncu
BENCH="python bench.py --profile-batch --frames 20 --batch 20 --no-graph --no-cpu-baseline"
# (1) launch list of the batch, EVERY kernel (no name filter: ATen kernels would show up here)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled \
    --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list exit $?"
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.txt 2>&1
# (2) full metric set of every launch of the batch
timeout 1500 ncu --profile-from-start off --set full --clock-control none -o /tmp/ncu/${TAG}_all $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture exit $?"
python tools/ncu_extract.py /tmp/ncu/${TAG}_all.ncu-rep gpurun_out/${TAG}_all_kernels_ncu.csv
# (3) one residual-block conv with source correlation, kept as a report
if [ "${2:-}" != "quick" ]; then
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:conv_halo_pair_kernel<256" -s 4 -c 1 \
      -o gpurun_out/${TAG}_conv_res $BENCH > gpurun_out/${TAG}_ncu_conv1.log 2>&1
fi
ls -la gpurun_out | tail -12

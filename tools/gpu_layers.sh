#!/bin/bash
# Run ON THE GPU BOX (via gpurun): conv + generator parity, then the per-launch conv timing table of one 50-frame batch
# for each precision given as arguments (default: fp16x2).  Output: gpurun_out/layers_<precision>.json
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_generator_gpu.py -x -q -m gpu 2>&1 | tail -5
for pr in "${@:-fp16x2}"; do
    python bench.py --no-cpu-baseline --precision "$pr" --steps 2 --warmup 2 --dump-layers "gpurun_out/layers_$pr.json" 2>&1 \
        | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$pr', 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'conv_ms', round(r['conv_ms_per_batch'],2), 'batch_ms', round(r['batch_ms_ungraphed'],2), 'frac', round(r['frac'],3), 'exec', round(r['executed_tflops']))"
done

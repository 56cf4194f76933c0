#!/bin/sh
for v in v0 v1 v2 v3 v5; do
  for cfg in "--interp Lanczos4" "--interp Bicubic" "--config 3 --interp Lanczos4"; do
    GF_CUDA_LIB=$PWD/gyroflow_b200/alt/libgf_$v.so timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline $cfg > /tmp/o.json 2>/tmp/o.err
    python -c "
import json; d=json.load(open('/tmp/o.json')); print('$v', '$cfg', round(d['value'],1))" 2>&1 | tail -1
  done
done

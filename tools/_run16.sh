#!/bin/sh
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "high_order or bicubic or ewa or lanczos or planes" 2>&1 | tail -2
for v in main b c d; do
  if [ $v = main ]; then unset GF_CUDA_LIB; else export GF_CUDA_LIB=$PWD/gyroflow_b200/alt/libgf_$v.so; fi
  for cfg in "--interp Lanczos4" "--interp Bicubic" "--config 3 --interp Lanczos4" "--config 3 --interp Bicubic"; do
    timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline $cfg > /tmp/o.json 2>/tmp/o.err
    python -c "
import json; d=json.load(open('/tmp/o.json')); print('$v', '$cfg', round(d['value'],1))" 2>&1 | tail -1
  done
done

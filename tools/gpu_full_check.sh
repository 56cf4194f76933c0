#!/bin/sh
# Runs ON THE GPU BOX (gpurun -- 'sh tools/gpu_full_check.sh'): the complete GPU suite, smoke(), compute-sanitizer, the high-order
# resampler side measurements and the default bench.  Everything lands in gpurun_out/.
timeout 2000 python -m pytest tests -m gpu -x -q > gpurun_out/full_pytest.log 2>&1; tail -3 gpurun_out/full_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
sh tools/sanitize_gpu.sh 2>&1 | grep -E "SUMMARY|done"
for i in Lanczos4 Bicubic "EWA: Robidoux"; do
  timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --interp "$i" > gpurun_out/full_interp.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/full_interp.json')); print('$i', round(d['value'],1))"
done
timeout 600 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; tail -c 1200 gpurun_out/full_bench.json

#!/bin/sh
timeout 2000 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_pytest.log 2>&1; tail -4 gpurun_out/r02m_pytest.log
for i in Lanczos4 Bicubic; do timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --interp $i > gpurun_out/r02m_$i.json 2> gpurun_out/r02m_$i.err; python - <<PY
import json; d=json.load(open("gpurun_out/r02m_$i.json")); print("$i", d["value"], d["ms_per_step"])
PY
done
timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --config 3 --interp Lanczos4 > gpurun_out/r02m_cfg3_lanczos.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02m_cfg3_lanczos.json')); print('cfg3 lanczos', d['value'])"
timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --config 3 --interp Bicubic > gpurun_out/r02m_cfg3_bicubic.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02m_cfg3_bicubic.json')); print('cfg3 bicubic', d['value'])"
sh tools/sanitize_gpu.sh

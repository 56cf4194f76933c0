#!/bin/sh
timeout 2000 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; tail -4 gpurun_out/r02l_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for i in Lanczos4 Bicubic; do timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --interp $i > gpurun_out/r02l_$i.json 2> gpurun_out/r02l_$i.err; python - <<PY
import json; d=json.load(open("gpurun_out/r02l_$i.json")); print("$i", d["value"], d["ms_per_step"], d.get("roofline",{}).get("launch_us"))
PY
done
timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --config 3 --interp Lanczos4 > gpurun_out/r02l_cfg3_lanczos.json 2>/dev/null; tail -c 600 gpurun_out/r02l_cfg3_lanczos.json
timeout 600 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; tail -c 1800 gpurun_out/r02l_bench.json

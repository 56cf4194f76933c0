#!/bin/sh
timeout 2000 python -m pytest tests -m gpu -x -q > gpurun_out/r02n_pytest.log 2>&1; tail -3 gpurun_out/r02n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
sh tools/sanitize_gpu.sh 2>&1 | grep -E "SUMMARY|done"
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
run() { name=$1; shift
    timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/lens_$name.json 2> gpurun_out/lens_$name.err
    timeout 300 ncu --metrics $M --clock-control none -k regex:warp_kernel -s 30 -c 1 --csv --log-file gpurun_out/lens_$name.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > /dev/null 2>&1
    python -c "
import json; d=json.load(open('gpurun_out/lens_$name.json')); print('$name', round(d['value'],1))"
}
run lanczos4 --interp Lanczos4
run bicubic --interp Bicubic
for i in "EWA: Robidoux"; do timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --interp "$i" > gpurun_out/r02n_ewa.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02n_ewa.json')); print('ewa robidoux', round(d['value'],1))"; done
timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --config 3 --interp Lanczos4 > gpurun_out/r02n_cfg3_lanczos.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02n_cfg3_lanczos.json')); print('cfg3 lanczos', round(d['value'],1))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shade_from_coords -s 10 -c 1 -o /tmp/r02n_lanczos4_shade -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --interp Lanczos4 > /dev/null 2>&1
python tools/ncu_summary.py /tmp/r02n_lanczos4_shade.ncu-rep 8294400 > gpurun_out/r02n_lanczos4_shade_summary.txt 2>&1; head -20 gpurun_out/r02n_lanczos4_shade_summary.txt | cut -c1-120
timeout 600 python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; tail -c 1500 gpurun_out/r02n_bench.json

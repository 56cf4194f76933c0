set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log; tail -4 gpurun_out/r02h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; tail -2 gpurun_out/r02h_smoke.log
timeout 600 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -c 1500 gpurun_out/r02h_bench.json; tail -3 gpurun_out/r02h_bench.err
timeout 300 python tools/filter_selftest.py > gpurun_out/r02_filter_selftest.txt 2>&1; cat gpurun_out/r02_filter_selftest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02h_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_kernel_x2 -s 40 -c 1 -o gpurun_out/r02h_head python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02h_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shade_from_coords -s 10 -c 1 -o gpurun_out/r02h_lanczos python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --interp Lanczos4 > gpurun_out/r02h_ncu_l4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_kernel -s 10 -c 1 -o gpurun_out/r02h_cfg4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --config 4 > gpurun_out/r02h_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_kernel -s 10 -c 1 -o gpurun_out/r02h_cfg3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --config 3 > gpurun_out/r02h_ncu_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep

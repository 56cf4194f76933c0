set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -30 gpurun_out/r02c_pytest.log
timeout 600 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 3500 gpurun_out/r02c_bench.json; tail -20 gpurun_out/r02c_bench.err
GF_DISABLE_FILTER=1 timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r02c_bench_nofilter.json 2>&1; tail -c 1500 gpurun_out/r02c_bench_nofilter.json
timeout 300 python tools/pcie_bidir.py > gpurun_out/r02c_pcie.json 2>&1; cat gpurun_out/r02c_pcie.json
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/bench_tex_gather.cu -o /tmp/bench_tex_gather && timeout 120 /tmp/bench_tex_gather > gpurun_out/r02c_tex_ab.json 2>&1; cat gpurun_out/r02c_tex_ab.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_kernel_x2 -s 40 -c 2 -o gpurun_out/r02c_head python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02c_ncu.log 2>&1

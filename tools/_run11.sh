set -x
timeout 1500 python -m pytest tests -m gpu -x -q -k "fisheye_filter_sweep or host_register or thousand_frame or overlays" > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log; tail -15 gpurun_out/r02k_pytest.log
timeout 600 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -c 1800 gpurun_out/r02k_bench.json; tail -3 gpurun_out/r02k_bench.err

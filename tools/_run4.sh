set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -15 gpurun_out/r02d_pytest.log
timeout 600 python bench.py > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; tail -c 3000 gpurun_out/r02d_bench.json; tail -5 gpurun_out/r02d_bench.err
GF_DISABLE_FILTER=1 timeout 600 python bench.py --no-e2e > gpurun_out/r02d_bench_nofilter.json 2>&1; tail -c 1200 gpurun_out/r02d_bench_nofilter.json
sh tools/sanitize_gpu.sh
sh tools/lens_table_gpu.sh

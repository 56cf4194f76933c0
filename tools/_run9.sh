set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log; tail -4 gpurun_out/r02i_pytest.log
timeout 600 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -c 1200 gpurun_out/r02i_bench.json; tail -3 gpurun_out/r02i_bench.err
timeout 600 python bench.py --lens sony --no-e2e --no-cpu-baseline > gpurun_out/r02i_bench_sony.json 2>&1; tail -c 400 gpurun_out/r02i_bench_sony.json

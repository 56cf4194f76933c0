set -x
export GF_RUN_EXHAUSTIVE=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -30 gpurun_out/r02b_pytest.log
timeout 600 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 4000 gpurun_out/r02b_bench.json; tail -20 gpurun_out/r02b_bench.err

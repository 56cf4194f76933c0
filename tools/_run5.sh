set -x
timeout 600 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -c 3000 gpurun_out/r02e_bench.json; tail -5 gpurun_out/r02e_bench.err
GF_DISABLE_FILTER=1 timeout 600 python bench.py --no-e2e > gpurun_out/r02e_bench_nofilter.json 2> gpurun_out/r02e_bench_nofilter.err; tail -c 1500 gpurun_out/r02e_bench_nofilter.json; tail -3 gpurun_out/r02e_bench_nofilter.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_bench_ref.json 2>&1; tail -c 600 gpurun_out/r02e_bench_ref.json
timeout 900 python -m pytest tests -m gpu -x -q -k "filter or variants or queue or cfg2" > gpurun_out/r02e_pytest.log 2>&1; tail -5 gpurun_out/r02e_pytest.log

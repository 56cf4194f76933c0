set -x
for mb in 5 4; do
  GF_CUDA_LIB=$PWD/gyroflow_b200/libgyroflow_cuda_mb$mb.so timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r02j_bench_mb$mb.json 2>&1; tail -c 700 gpurun_out/r02j_bench_mb$mb.json
done
timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r02j_bench_mb6.json 2>&1; tail -c 700 gpurun_out/r02j_bench_mb6.json

set -x
nvidia-smi topo -m > gpurun_out/r02g_topo8.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02g_bench_8gpu.json 2> gpurun_out/r02g_bench_8gpu.err; tail -c 2500 gpurun_out/r02g_bench_8gpu.json; tail -5 gpurun_out/r02g_bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02g_bench_4gpu.json 2> gpurun_out/r02g_bench_4gpu.err; tail -c 1500 gpurun_out/r02g_bench_4gpu.json

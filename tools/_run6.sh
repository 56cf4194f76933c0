set -x
nvidia-smi topo -m > gpurun_out/r02f_topo2.txt 2>&1
timeout 600 python -m pytest tests/test_render_queue.py -m gpu -x -q > gpurun_out/r02f_pytest_2gpu.log 2>&1; tail -5 gpurun_out/r02f_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02f_bench_2gpu.json 2> gpurun_out/r02f_bench_2gpu.err; tail -c 2500 gpurun_out/r02f_bench_2gpu.json; tail -5 gpurun_out/r02f_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02f_ref_2gpu.json 2>&1; tail -c 500 gpurun_out/r02f_ref_2gpu.json

#!/bin/sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shade_from_coords -s 20 -c 1 -o gpurun_out/r02l_lanczos_shade -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --interp Lanczos4 > /dev/null 2>&1
ls -la gpurun_out/r02l_lanczos_shade.ncu-rep

#!/bin/sh
# Runs ON THE GPU BOX: for every lens model (and the fisheye x digital-lens pairs) of BASELINE config 2's frame (3840x2160 RGBA8, rolling
# shutter on, bilinear): the device-resident frame rate (bench.py side measurement) and one light ncu pass over the dominant kernel.
# Output: gpurun_out/lens_<name>.json / .csv, turned into profiles/LENS_TABLE.md by tools/lens_table.py here.
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
run() {   # name, bench args
    name=$1; shift
    timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/lens_$name.json 2> gpurun_out/lens_$name.err
    timeout 300 ncu --metrics $M --clock-control none -k regex:warp_kernel -s 30 -c 1 --csv --log-file gpurun_out/lens_$name.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > /dev/null 2>&1
}
run opencv_fisheye --legacy
for l in opencv_standard poly3 poly5 ptlens insta360 sony generic_polynomial gopro; do run $l --lens $l; done
for d in gopro_superview gopro6_superview gopro_hyperview digital_stretch; do run opencv_fisheye+$d --digital $d; done
run gopro+gopro_warp --lens gopro --digital gopro_warp
run cfg3_luma16_superview --config 3
run cfg4_r32f_sony_ibis_mesh --config 4
run cfg4_fused_4_planes --config 4 --planes 4
run lanczos4 --interp Lanczos4
run bicubic --interp Bicubic
ls gpurun_out/lens_* | wc -l

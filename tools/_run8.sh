set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log; tail -4 gpurun_out/r02h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; tail -2 gpurun_out/r02h_smoke.log
timeout 600 python bench.py > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -c 1500 gpurun_out/r02h_bench.json; tail -3 gpurun_out/r02h_bench.err
timeout 600 python bench.py --depth 8 --no-e2e --no-cpu-baseline > gpurun_out/r02h_bench_depth8.json 2>&1; tail -c 600 gpurun_out/r02h_bench_depth8.json
timeout 300 python tools/filter_selftest.py > gpurun_out/r02_filter_selftest.txt 2>&1; cat gpurun_out/r02_filter_selftest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02h_launches.log 2>&1
cap() {  # name, kernel regex, skip, pixels, bench args...
  name=$1; rx=$2; skip=$3; px=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/$name python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${name}_ncu.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep $px > gpurun_out/${name}_summary.txt 2>&1
}
cap r02h_x2_filtered_fisheye_rgba8 warp_kernel_x2 40 8294400
cp /tmp/r02h_x2_filtered_fisheye_rgba8.ncu-rep gpurun_out/
cap r02h_lanczos4_shade shade_from_coords 10 8294400 --interp Lanczos4
cap r02h_cfg4_r32f_sony_ibis_mesh warp_kernel 10 8294400 --config 4
cap r02h_cfg3_luma16_superview warp_kernel 10 33177600 --config 3
cap r02h_gopro_packed warp_kernel_x2 10 8294400 --lens gopro
du -sh gpurun_out

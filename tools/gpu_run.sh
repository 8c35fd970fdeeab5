mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ll_plain.json 2>gpurun_out/ll_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:solver_kernel|warp_|gauss_|resample_|gray_|minmax_|setup_|interleave_|poison_|demosaic_|upsample2_|remosaick" -c 300 --csv --log-file gpurun_out/launches_r02e.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ll_ncu.log 2>&1
echo "rc $?"; wc -l gpurun_out/launches_r02e.csv

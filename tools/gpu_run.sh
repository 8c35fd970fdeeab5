mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:solver_kernel|warp_|gauss_|resample_|gray_|minmax_|setup_|interleave_|poison_|demosaic_|upsample2_|remosaick" -c 300 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo ncu_rc=$?
python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench.json 2>/dev/null

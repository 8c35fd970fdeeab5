mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v Warning | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/p_base.json 2> gpurun_out/p_base.err
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:solver_kernel|warp_|gauss_|resample_|gray_|minmax_|setup_|interleave_|demosaic_|upsample2_|remosaick" -c 300 --csv --log-file gpurun_out/launches_r1p.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1

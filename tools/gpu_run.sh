mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/x_long.json 2> gpurun_out/x_long.err

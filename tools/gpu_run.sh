mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v Warning | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s_base.json 2> gpurun_out/s_base.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s_base2.json 2> gpurun_out/s_base2.err

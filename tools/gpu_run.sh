mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -8
python profiles/bench_warp.py 2>&1 | grep shape | cut -c1-200
python bench.py --steps 5 --warmup 3 > gpurun_out/d_base.json 2> gpurun_out/d_base.err
tail -c 1500 gpurun_out/d_base.json

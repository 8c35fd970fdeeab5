mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
RVDD_FUSE=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline_configs.py -m gpu -x -q 2>&1 | tail -2
for r in 1 2; do
RVDD_FUSE=0 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_never_$r.json 2>/dev/null
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_auto_$r.json 2>/dev/null
done
for n in clean iso12800; do
RVDD_FUSE=0 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --noise $n > gpurun_out/r2t_never_$n.json 2>/dev/null
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --noise $n > gpurun_out/r2t_auto_$n.json 2>/dev/null
done

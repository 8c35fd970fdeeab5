mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -6
for px in 0 3072 1536 768; do
echo "== RVDD_PX_PER_CTA=$px"
RVDD_PX_PER_CTA=$px timeout 120 python tools/time_single.py 2>&1 | tail -4 | cut -c1-200
RVDD_PX_PER_CTA=$px timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_px${px}.json 2>/dev/null
done

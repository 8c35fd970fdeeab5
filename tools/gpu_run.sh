mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -s -k "not warp and not demosaic" 2>&1 | grep -E "passed|failed|hypot:|Error|error" | head
for r in 1 2 3; do
for v in prev base; do
  if [ $v = base ]; then L=rvdd-release_b200/lib/libBridge.so; else L=rvdd-release_b200/lib/libBridge_$v.so; fi
  RVDD_BRIDGE_LIB=$L python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_${v}_$r.json 2>/dev/null
done; done
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --noise clean > gpurun_out/r2o_clean.json 2>/dev/null

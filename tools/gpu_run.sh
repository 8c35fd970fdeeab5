mkdir -p gpurun_out
for r in 1 2 3; do
for v in prev base; do
  if [ $v = base ]; then L=rvdd-release_b200/lib/libBridge.so; else L=rvdd-release_b200/lib/libBridge_$v.so; fi
  RVDD_BRIDGE_LIB=$L python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_${v}_$r.json 2>/dev/null
done; done
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -x -q -k "flow or instantiations or auto" 2>&1 | tail -1

mkdir -p gpurun_out
for r in 1 2; do
for v in base f2call; do
  if [ $v = base ]; then L=rvdd-release_b200/lib/libBridge.so; else L=rvdd-release_b200/lib/libBridge_$v.so; fi
  RVDD_BRIDGE_LIB=$L python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_${v}_$r.json 2>/dev/null
done; done
for v in base f2call; do
  if [ $v = base ]; then L=rvdd-release_b200/lib/libBridge.so; else L=rvdd-release_b200/lib/libBridge_$v.so; fi
  RVDD_FUSE=1 RVDD_BRIDGE_LIB=$L python bench.py --steps 8 --warmup 3 --no-cpu-baseline --noise clean > gpurun_out/r2x_${v}_clean_forced.json 2>/dev/null
done
RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge_f2call.so python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -k "instantiations" 2>&1 | tail -1

mkdir -p gpurun_out
for r in 1 2; do
for v in base cq hz cqhz; do
  if [ $v = base ]; then L=rvdd-release_b200/lib/libBridge.so; else L=rvdd-release_b200/lib/libBridge_$v.so; fi
  RVDD_BRIDGE_LIB=$L python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/z_${v}_$r.json 2>/dev/null
done; done
for v in cq hz cqhz; do RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flow or selftest" 2>&1 | tail -1; done

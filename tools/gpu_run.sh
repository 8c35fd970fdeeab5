mkdir -p gpurun_out
for v in wb40 wb36; do RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge_$v.so python profiles/bench_warp.py 2>&1 | grep HWC; done
for r in 1 2; do
for v in prev base; do
  if [ $v = base ]; then L=rvdd-release_b200/lib/libBridge.so; else L=rvdd-release_b200/lib/libBridge_$v.so; fi
  RVDD_BRIDGE_LIB=$L python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_${v}_$r.json 2>/dev/null
done; done

mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for r in 1 2; do
for v in base hf32; do
  if [ $v = base ]; then L=rvdd-release_b200/lib/libBridge.so; else L=rvdd-release_b200/lib/libBridge_$v.so; fi
  RVDD_BRIDGE_LIB=$L python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_${v}_$r.json 2>/dev/null
done; done
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --noise clean > gpurun_out/r2b_clean.json 2>/dev/null
python tools/prof_solver.py 29 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:solver_kernel -s 2 -c 1 -o gpurun_out/solver_r02a python tools/prof_solver.py 29 > gpurun_out/ncu.log 2>&1
echo ncu_rc=$?

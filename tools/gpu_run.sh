mkdir -p gpurun_out
for r in 1 2; do
for v in s2 s3nf; do
RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge_$v.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_${v}_$r.json 2>/dev/null
done
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_fused_$r.json 2>/dev/null
done
python tools/prof_solver.py 29 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:solver_kernel -s 2 -c 1 -o gpurun_out/solver_r02b_fused python tools/prof_solver.py 29 > gpurun_out/ncu.log 2>&1
echo ncu_rc=$?

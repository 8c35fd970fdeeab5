mkdir -p gpurun_out
python -m pytest tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -2
for m in never auto always; do echo "== $m"; python tools/time_single.py $m 2>&1 | tail -4 | cut -c1-160; done
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_auto.json 2>/dev/null
python tools/prof_solver.py 29 iso3200 always > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:solver_kernel -s 2 -c 1 -o gpurun_out/solver_r02d_fused python tools/prof_solver.py 29 iso3200 always > gpurun_out/ncu.log 2>&1
echo ncu_rc=$?

mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2>/dev/null
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --noise clean > gpurun_out/r2p_clean.json 2>/dev/null
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --noise iso12800 > gpurun_out/r2p_iso12800.json 2>/dev/null
python tools/time_single.py 2>&1 | tail -4 | cut -c1-220
python tools/prof_solver.py 29 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:solver_kernel -s 2 -c 1 -o gpurun_out/solver_r02c python tools/prof_solver.py 29 > gpurun_out/ncu.log 2>&1
echo ncu_rc=$?

mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py > gpurun_out/h_default.json 2> gpurun_out/h_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/h_ref.json 2> gpurun_out/h_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solver_kernel -s 3 -c 1 -o gpurun_out/prof_solver_r1i -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_solver.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_hwc4 -s 1 -c 1 -o gpurun_out/prof_warp_hwc4_r1a -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_hwc4.log 2>&1
echo ok

mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -3
python bench.py > gpurun_out/q_default.json 2> gpurun_out/q_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/q_ref.json 2> gpurun_out/q_ref.err
ncu --set full --clock-control none --import-source on -k regex:solver_kernel -s 3 -c 1 -o gpurun_out/prof_solver_r1q -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_solver.log 2>&1
python profiles/bench_warp.py > gpurun_out/stage_kernels.jsonl 2>gpurun_out/stage_kernels.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log

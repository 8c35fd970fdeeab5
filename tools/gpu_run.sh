mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c_t2.json 2> gpurun_out/c_t2.err
for v in 1 4; do RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge_t$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c_t$v.json 2> gpurun_out/c_t$v.err; done

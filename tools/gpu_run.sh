mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge_p1.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flow" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g_p2.json 2> gpurun_out/g_p2.err
RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge_p1.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g_p1.json 2> gpurun_out/g_p1.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g_p2b.json 2> gpurun_out/g_p2b.err
RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge_p1.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g_p1b.json 2> gpurun_out/g_p1b.err

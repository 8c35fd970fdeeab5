mkdir -p gpurun_out
for v in "" _f2dhw; do
  RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge$v.so python tools/time_wc.py 2>&1 | tail -1 | tee -a gpurun_out/f2d_wc.txt
done
python -m pytest tests/test_gpu_dropin.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3

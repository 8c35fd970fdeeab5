mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_warp or dropin or precompute" 2>&1 | tail -2
python tools/time_dropin.py 2>&1 | tail -2

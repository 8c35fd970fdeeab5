mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v Warning | grep -B30 "Error\|error" | tail -50

python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v Warning | tail -2
timeout 600 python tools/time_sizes.py 2>&1 | tail -3

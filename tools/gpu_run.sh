mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -6

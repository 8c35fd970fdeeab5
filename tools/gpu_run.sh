mkdir -p gpurun_out
for i in 1 2 3; do python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -1; done

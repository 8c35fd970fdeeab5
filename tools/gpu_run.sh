mkdir -p gpurun_out
timeout 600 python tools/time_sizes.py 2>&1 | tail -4

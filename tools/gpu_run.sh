mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline_configs.py -m gpu -x -q -s -k "c3 or c5" 2>&1 | tail -25
python -m pytest tests/test_gpu_dropin.py -m gpu -x -q -k "fused" 2>&1 | tail -5

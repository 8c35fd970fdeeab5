python profiles/bench_warp.py --demosaic-only | cut -c1-140
python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_pipeline_configs.py -m gpu -x -q 2>&1 | tail -2

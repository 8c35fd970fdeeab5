python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -m gpu -x -q -k "demosaic or variants or pipeline or aligner" 2>&1 | tail -2
python profiles/bench_warp.py --demosaic-only | cut -c1-150

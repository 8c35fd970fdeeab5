mkdir -p gpurun_out
python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | grep -v Warning | tail -3
python profiles/bench_align.py > gpurun_out/align.jsonl 2> gpurun_out/align.err; cat gpurun_out/align.jsonl; tail -3 gpurun_out/align.err

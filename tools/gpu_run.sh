mkdir -p gpurun_out
python profiles/bench_warp.py --hwc-only > gpurun_out/w5_2rows.jsonl 2>gpurun_out/w5.err
RVDD_WARP_HWC_1PX=1 python profiles/bench_warp.py --hwc-only > gpurun_out/w5_1px.jsonl 2>>gpurun_out/w5.err
cat gpurun_out/w5_2rows.jsonl gpurun_out/w5_1px.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/w5_pytest.log 2>&1; echo "pytest rc $?"; tail -1 gpurun_out/w5_pytest.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/w5_bench.json 2>gpurun_out/w5_bench.err; cat gpurun_out/w5_bench.json

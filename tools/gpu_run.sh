mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== auto"; python tools/time_single.py auto 2>&1 | tail -4 | cut -c1-100
python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench.json 2>/dev/null
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --noise clean > gpurun_out/r2v_clean.json 2>/dev/null
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --noise iso12800 > gpurun_out/r2v_iso12800.json 2>/dev/null
python profiles/bench_precompute.py --seqs-per-gpu 12 > gpurun_out/r2v_c4_1gpu.json 2>/dev/null

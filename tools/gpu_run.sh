mkdir -p gpurun_out
python bench.py --noise clean --no-cpu-baseline > gpurun_out/fin_clean.json 2>gpurun_out/fin_clean.err; echo "rc $?"
python bench.py --noise iso12800 --no-cpu-baseline > gpurun_out/fin_12800.json 2>gpurun_out/fin_12800.err; echo "rc $?"
python tools/show.py gpurun_out/fin_clean.json | head -2; python tools/show.py gpurun_out/fin_12800.json | head -2

mkdir -p gpurun_out
python bench.py --noise clean > gpurun_out/v_clean.json 2> gpurun_out/v_clean.err
python bench.py --noise iso12800 > gpurun_out/v_iso12800.json 2> gpurun_out/v_iso12800.err
python bench.py --noise iso3200 --no-cpu-baseline > gpurun_out/v_iso3200.json 2> gpurun_out/v_iso3200.err

mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --groups 14 > gpurun_out/m_g14.json 2> gpurun_out/m_g14.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --groups 29 > gpurun_out/m_g29.json 2> gpurun_out/m_g29.err

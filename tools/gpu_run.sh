mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -3
python bench.py > gpurun_out/t_default.json 2> gpurun_out/t_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/t_ref.json 2> gpurun_out/t_ref.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

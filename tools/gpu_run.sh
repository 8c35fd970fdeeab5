mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "precompute" 2>&1 | tail -3
df -h /dev/shm /tmp | tail -3; nproc; free -g | head -2
python profiles/bench_precompute.py --seqs-per-gpu 12 > gpurun_out/c4_1gpu_shm.json 2> gpurun_out/c4_1gpu_shm.err; tail -3 gpurun_out/c4_1gpu_shm.err
python profiles/bench_precompute.py --seqs-per-gpu 12 --readers 8 --writers 8 > gpurun_out/c4_1gpu_shm_r8w8.json 2>/dev/null
python profiles/bench_precompute.py --seqs-per-gpu 12 --readers 2 --writers 2 > gpurun_out/c4_1gpu_shm_r2w2.json 2>/dev/null
python profiles/bench_precompute.py --seqs-per-gpu 6 --root /tmp/rvdd_config4 > gpurun_out/c4_1gpu_disk.json 2>/dev/null
cat gpurun_out/c4_*.json

mkdir -p gpurun_out
python tools/time_sizes.py | tee gpurun_out/sizes_default.jsonl
RVDD_FUSE_MIN_PX=200000 python tools/time_sizes.py | tee gpurun_out/sizes_fuse200k.jsonl
RVDD_FUSE_MIN_PX=200000 RVDD_FUSE_MIN_ROWS=32 python tools/time_sizes.py | tee gpurun_out/sizes_fuse200k_r32.jsonl

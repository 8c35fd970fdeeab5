mkdir -p gpurun_out
python profiles/bench_precompute.py --seqs-per-gpu 12 2>gpurun_out/c4a.err | grep '^{' > gpurun_out/c4_default.json
RVDD_FUSE_MIN_PX=200000 python profiles/bench_precompute.py --seqs-per-gpu 12 2>gpurun_out/c4b.err | grep '^{' > gpurun_out/c4_fuse200k.json
python - <<'PY'
import json
for f in ("gpurun_out/c4_default.json", "gpurun_out/c4_fuse200k.json"):
    d = json.load(open(f))
    print(f, round(d["pairs_per_s_files_included"]), round(d["pairs_per_s_no_io_same_batches"]), d["files_bit_equal_to_direct_compute"])
PY

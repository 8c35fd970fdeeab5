# usage: bash tools/gpu_run8.sh N   (inside gpurun --gpus N): bench, config 4 through files
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2w_bench_${N}gpu.json 2> gpurun_out/r2w_bench_${N}gpu.err
$TR profiles/bench_precompute.py --seqs-per-gpu 24 > gpurun_out/r2w_c4_${N}gpu.json 2> gpurun_out/r2w_c4_${N}gpu.err
grep '^{' gpurun_out/r2w_c4_${N}gpu.json | cut -c1-500

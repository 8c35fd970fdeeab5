# usage: bash tools/gpu_run8.sh N [tag extra-args...]   (inside gpurun --gpus N)
N=${1:-8}; TAG=${2:-auto}; shift; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR profiles/bench_precompute.py --seqs-per-gpu 24 "$@" > gpurun_out/r2g_c4_${N}gpu_$TAG.json 2> gpurun_out/r2g_c4_${N}gpu_$TAG.err; tail -c 300 gpurun_out/r2g_c4_${N}gpu_$TAG.err
cat gpurun_out/r2g_c4_${N}gpu_$TAG.json

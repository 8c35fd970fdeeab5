mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/topo.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/e_8gpu.json 2> gpurun_out/e_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-numa-bind > gpurun_out/e_8gpu_nobind.json 2> gpurun_out/e_8gpu_nobind.err
grep -o '"e2e": {[^}]*}' gpurun_out/e_8gpu.json gpurun_out/e_8gpu_nobind.json

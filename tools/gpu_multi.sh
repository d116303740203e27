#!/bin/bash
# bench.py on N GPUs of one box, both arms, as the driver launches them: tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; cut -c1-300 gpurun_out/bench_${N}gpu.json; tail -3 gpurun_out/bench_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_${N}gpu_ref.json 2>> gpurun_out/bench_${N}gpu.err; cut -c1-300 gpurun_out/bench_${N}gpu_ref.json

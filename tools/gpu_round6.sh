mkdir -p gpurun_out
python tools/parity_report.py --out gpurun_out/parity.json 2>&1 | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2>> gpurun_out/bench_2gpu.err; cut -c1-300 gpurun_out/bench_2gpu_ref.json

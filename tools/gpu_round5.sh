mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
# the shuffle-exchange FFT experiment: parity first (a wrong kernel is not a measurement), then its time
FXB200_LIB=$PWD/feature-extractor_b200/lib/exp/libfxb200_shfl23.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "parity_against_oracle or golden" 2>&1 | tail -3
bash tools/gpu_ab.sh 4096 1024 2>&1 | tee gpurun_out/ab_4096.txt
bash tools/gpu_ab.sh 2048 1024 2>&1 | tee gpurun_out/ab_2048.txt
for c in 1 2 3 4 5 6 8; do FXB200_CHUNKS=$c python bench.py --no-cpu --no-e2e --no-c5 --no-rt --steps 4 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunks', $c, round(d['roofline']['kernel_ms'], 2))"; done | tee gpurun_out/chunks.txt

mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
for so in feature-extractor_b200/lib/exp/*.so; do
  echo "== parity of $so"
  FXB200_LIB=$PWD/$so timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "parity_against_oracle or golden" 2>&1 | tail -2
done
bash tools/gpu_ab.sh 4096 1024 2>&1 | tee gpurun_out/ab_4096.txt
bash tools/gpu_ab.sh 2048 1024 2>&1 | tee gpurun_out/ab_2048.txt

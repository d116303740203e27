mkdir -p gpurun_out
for so in feature-extractor_b200/lib/exp/*.so; do
  echo "== parity of $so"
  FXB200_LIB=$PWD/$so timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "parity_against_oracle or golden or config2 or ten_minute" 2>&1 | tail -4
done
bash tools/gpu_ab.sh 4096 1024 2>&1 | tee gpurun_out/ab_4096.txt
bash tools/gpu_ab.sh 2048 1024 2>&1 | tee gpurun_out/ab_2048.txt

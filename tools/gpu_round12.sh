mkdir -p gpurun_out
for so in feature-extractor_b200/lib/exp/libfxb200_occ.so; do
  echo "== parity of $so"
  FXB200_LIB=$PWD/$so timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "parity_against_oracle or golden or chunking" 2>&1 | tail -2
done
bash tools/gpu_ab.sh 2048 1024 2>&1 | tee gpurun_out/ab_2048.txt
bash tools/gpu_ab.sh 1024 512 2>&1 | tee gpurun_out/ab_1024.txt
bash tools/gpu_ab.sh 2048 512 2>&1 | tee gpurun_out/ab_2048_512.txt

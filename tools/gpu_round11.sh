mkdir -p gpurun_out
for g in 8 12 16 24 32 48; do FXB200_PIPE_GROUPS=$g python bench.py --no-cpu --no-c5 --no-rt --steps 3 --warmup 3 --e2e-steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('groups', $g, 'e2e', round(d['e2e']['value']/1e6,3), 'pcm16', round(d['e2e_pcm16']['value']/1e6,3), 'pcm ms', round(d['e2e_pcm16']['ms_per_step']['mean'],2))"; done | tee gpurun_out/pipe_groups.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k facade 2>&1 | tail -3

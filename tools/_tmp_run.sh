mkdir -p gpurun_out
export SDSP_B200_FFT_FUSED_TMA=1
timeout 150 python -m pytest tests/test_gpu_fft.py -x -q -k "all_sizes or large_frames or real_input or fused_65536" 2>&1 | tail -5
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-secondary --workload"
for w in fft65536_f32 fft32768_f32; do
  echo "== $w tma"; timeout 60 $B $w | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
  echo "== $w v3"; SDSP_B200_FFT_FUSED_TMA=0 timeout 60 $B $w | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
done

mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_fft.py -x -q -k "all_sizes or large_frames or real_input or fused_65536" 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-secondary --workload"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["roofline"]["frac"])'
for v in "" _ef; do
  export SDSP_B200_LIB=/root/repo/simpledsp_b200/lib$v/libsdsp_b200.so
  for w in fft65536_f32 fft32768_f32; do
    echo "== lib$v $w"; timeout 60 $B $w | python -c "$P"
  done
done
